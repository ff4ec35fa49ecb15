/* lines_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's line extraction: what slam_ros/main.cpp:37-71 (`mapping_cb`, SIMULATIONOFF
 * branch) does with one `mappingPoints` payload -- polar points, `LineExtraction` (slam_ros/lineFitting.cpp:640-702),
 * the final alfa += pi -- written as plain loops over arrays, operation for operation in the reference's order
 * (same operand order, same accumulation order, glibc sin/cos/atan2), so that it agrees BITWISE with the
 * deterministic instance of the reference built as oracle/_ref/libslamlines.so (tests/test_lines_oracle.py).
 *
 * The reference reads indeterminate memory on this path; the semantics fixed here (and in that build) are:
 *   - residual_error's `sum_dist` starts at 0                                   (lineFitting.cpp:109, 122)
 *   - Covariancia's C_x is the diagonal the code writes, zero elsewhere         (lineFitting.cpp:382, 416-420)
 *     and its angular entries are `1/12*1.5` == 0 (integer division), so the 2p refits that perturb the ANGLES
 *     (:422-441) multiply an exact zero and are not evaluated here.
 * Constants: PI is the truncated 3.14159265 of lineFitting.h:12 everywhere except main.cpp (M_PI).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may load this library.
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

const double PI = 3.14159265;                 /* lineFitting.h:12 */

struct PolarPoint { double alfa, r, weight, variance; };          /* simplifyPath.h:48-60 */
struct Line {                                                       /* simplifyPath.h:62-79 */
  double alfa, r, b, m;
  double C[4];
  double ia[2], ir[2];                                              /* lineInterval[0..1] */
};

/* line::line(alfa_deg, r), lineFitting.cpp:17-23 */
void make_line(Line& l, double alfa_deg, double r) {
  l.alfa = alfa_deg * (PI / 180);
  l.r = r;
  l.b = l.r / std::sin(l.alfa);
  l.m = -1 / std::tan(l.alfa);
}

/* lineFitting(vector<polar_point>&), lineFitting.cpp:267-304 */
Line fit(const PolarPoint* p, int n) {
  double sum_1 = 0, sum_2 = 0, sum_3 = 0, sum_4 = 0, sum_r = 0, wi = 0;
  for (int i = 0; i < n; i++) wi = wi + p[i].weight;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++)
      sum_1 = sum_1 + p[i].weight * p[j].weight * p[i].r * p[j].r * std::sin(p[i].alfa + p[j].alfa);
  for (int i = 0; i < n; i++)
    sum_2 = sum_2 + (p[i].weight - wi) * p[i].weight * p[i].r * p[i].r * std::sin(2 * p[i].alfa);
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++)
      sum_3 = sum_3 + p[i].weight * p[j].weight * p[i].r * p[j].r * std::cos(p[i].alfa + p[j].alfa);
  for (int i = 0; i < n; i++)
    sum_4 = sum_4 + (p[i].weight - wi) * p[i].weight * p[i].r * p[i].r * std::cos(2 * p[i].alfa);
  const double alfa = 0.5 * std::atan2((2 / wi) * sum_1 + (1 / wi) * sum_2, (2 / wi) * sum_3 + (1 / wi) * sum_4);
  for (int i = 0; i < n; i++) sum_r = sum_r + p[i].weight * p[i].r * std::cos(p[i].alfa - alfa);
  Line l;
  std::memset(&l, 0, sizeof l);
  make_line(l, alfa * 180 / PI, sum_r / wi);
  return l;
}

/* LineAlap, lineFitting.cpp:357-367 */
void canonical(double& alfa, double& r) {
  if (r < 0) {
    r = std::fabs(r);
    if (alfa < 0) alfa = PI + alfa; else alfa = -PI + alfa;
  }
}
double alfanorm(double a) {                    /* lineFitting.cpp:369-377 */
  if (a > PI) return a - 2 * PI;
  if (a < -PI) return a + 2 * PI;
  return a;
}

/* Covariancia, lineFitting.cpp:379-450 (range perturbations only; see the header) */
void covariance(std::vector<PolarPoint> pts, double C[4]) {
  const int n = (int)pts.size();
  Line base = fit(pts.data(), n);
  canonical(base.alfa, base.r);
  if (base.alfa < 0) base.alfa = base.alfa + 2 * PI;
  const double eps = 0.000001;
  std::vector<double> F0(n), F1(n);
  for (int i = 0; i < n; i++) {
    const double repo = pts[i].r;
    pts[i].r = pts[i].r + eps;
    Line e = fit(pts.data(), n);
    canonical(e.alfa, e.r);
    const double alfa = (e.alfa < 0) ? e.alfa + 2 * PI : e.alfa;
    F0[i] = alfanorm(alfa - base.alfa) / eps;
    F1[i] = (e.r - base.r) / eps;
    pts[i].r = repo;
  }
  /* (F C_x) F^T through the two reference-order dgemm calls (:443-444): C_x diagonal, zero-skip on the products */
  double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
  for (int k = 0; k < n; k++) {
    const double cx = pts[k].variance * pts[k].variance * 1.5;
    const double f0 = 0.0 + F0[k] * cx, f1 = 0.0 + F1[k] * cx;      /* F C_x: one non-zero product per entry */
    if (f0 != 0.0) { c00 += f0 * F0[k]; c01 += f0 * F1[k]; }
    if (f1 != 0.0) { c10 += f1 * F0[k]; c11 += f1 * F1[k]; }
  }
  (void)c01; (void)c10;
  C[0] = c00; C[1] = 0; C[2] = 0; C[3] = c11;                       /* :446-448 */
}

struct V2 { double x, y; };
double v2len(V2 a) {                           /* Vec2::Lenght, vec2.cpp:26-34 */
  const double t = std::pow(a.x, 2) + std::pow(a.y, 2);
  return (t > 0) ? std::sqrt(std::pow(a.x, 2) + std::pow(a.y, 2)) : 0;
}
V2 to_xy(double alfa, double r) { V2 v; v.x = std::cos(alfa) * r; v.y = std::sin(alfa) * r; return v; }

/* FirstPoint / EndPoint, simplifyPath.cpp:59-105: the foot, on the fitted line, of the segment's first / last point */
void end_point(const PolarPoint* p, int n, const Line& l, bool first_end, double& oa, double& orr) {
  const int e = first_end ? 0 : n - 1;
  if (n < 4) { oa = p[e].alfa; orr = p[e].r; return; }
  const double fr = l.r / (std::cos(p[n / 2].alfa - l.alfa));
  const double er = l.r / (std::cos(p[e].alfa - l.alfa));
  const double fa = p[n / 2].alfa, ea = p[e].alfa;
  const V2 P = to_xy(p[e].alfa, p[e].r);
  const V2 vf = to_xy(fa, fr), ve = to_xy(ea, er);
  V2 FE; FE.x = ve.x - vf.x; FE.y = ve.y - vf.y;
  V2 FP; FP.x = P.x - vf.x; FP.y = P.y - vf.y;
  const double lenFE = v2len(FE);
  V2 nrm; nrm.x = FE.x / lenFE; nrm.y = FE.y / lenFE;               /* Vec2::Norm */
  const double skal = FE.x * FP.x + FE.y * FP.y;                     /* SkalarCos, vec2.cpp:98-102 */
  const double cs = skal / (v2len(FE) * v2len(FP));
  const double sc = cs * v2len(FP);
  V2 N; N.x = vf.x + nrm.x * sc; N.y = vf.y + nrm.y * sc;
  orr = std::sqrt(N.x * N.x + N.y * N.y);                            /* descart2polar(Vec2), lineFitting.cpp:150-155 */
  oa = std::atan2(N.y, N.x);
}

/* simplifyPath::simplifyWithRDP, simplifyPath.cpp:108-177 */
void rdp(const PolarPoint* p, int n, std::vector<Line>& out) {
  if (n < 2) return;
  double sum_di = 0, sum_var = 0;
  Line sl = fit(p, n);
  for (int i = 0; i < n; i++)
    sum_di = sum_di + std::fabs(std::cos(p[i].alfa - sl.alfa)) * 2 * (p[i].variance) / (std::sqrt(2 * PI));
  for (int i = 0; i < n; i++)
    sum_var = sum_var + std::cos(p[i].alfa - sl.alfa) * std::cos(p[i].alfa - sl.alfa) * p[i].variance * p[i].variance * ((PI - 2) / PI);
  sum_var = std::sqrt(sum_var);
  std::vector<V2> c(n);
  for (int i = 0; i < n; i++) c[i] = to_xy(p[i].alfa, p[i].r);
  /* residual_error, lineFitting.cpp:95-125: distances to the chord of the FITTED line between the first and last abscissa */
  double t;
  {
    V2 f, l;
    f.x = c[0].x; f.y = sl.b + c[0].x * sl.m;
    l.x = c[n - 1].x; l.y = sl.b + c[n - 1].x * sl.m;
    V2 d; d.x = l.x - f.x; d.y = l.y - f.y;
    double sum_dist = 0;
    for (int i = 1; i < n; i++) {
      V2 pp; pp.x = c[i].x - f.x; pp.y = c[i].y - f.y;
      const double dist = std::fabs(pp.x * d.y - d.x * pp.y) / std::sqrt(d.x * d.x + d.y * d.y);
      sum_dist = sum_dist + dist;
    }
    t = sum_dist;
  }
  /* findMaximumDistance(points), simplifyPath.cpp:36-57: farthest from the chord first point -> last point */
  int index = 0;
  {
    V2 d; d.x = c[n - 1].x - c[0].x; d.y = c[n - 1].y - c[0].y;
    double md = -1;
    for (int i = 1; i < n; i++) {
      V2 pp; pp.x = c[i].x - c[0].x; pp.y = c[i].y - c[0].y;
      const double dist = std::fabs(pp.x * d.y - d.x * pp.y) / std::sqrt(d.x * d.x + d.y * d.y);
      if (dist > md) { md = dist; index = i; }
    }
  }
  if (t > sum_di + sum_var * 3) {
    if (index <= 0 || index >= n) return;      /* the reference would recurse forever here (NaN distances) */
    rdp(p, index, out);
    rdp(p + index, n - index, out);
    return;
  }
  covariance(std::vector<PolarPoint>(p, p + n), sl.C);
  end_point(p, n, sl, true, sl.ia[0], sl.ir[0]);
  end_point(p, n, sl, false, sl.ia[1], sl.ir[1]);
  for (int k = 0; k < 2; k++) sl.ir[k] = sl.r / (std::cos(sl.ia[k] - sl.alfa));   /* line::SetEndPoints, lineFitting.cpp:44-51 */
  for (int k = 0; k < 2; k++) sl.ia[k] = sl.ia[k] + PI;
  for (int k = 0; k < 2; k++) sl.ia[k] = sl.ia[k] > PI ? sl.ia[k] - 2.0 * PI : sl.ia[k];
  out.push_back(sl);
}

/* segmentation, lineFitting.cpp:541-584: split where neighbours are more than half a metre apart */
void segmentation(const std::vector<PolarPoint>& p, std::vector<int>& split) {
  const int len = (int)p.size() - 1;
  for (int i = 0; i < len; i++) {
    const double dist = std::sqrt(std::pow(p[i].r, 2) + std::pow(p[i + 1].r, 2) - 2 * p[i].r * p[i + 1].r * std::cos(p[i + 1].alfa - p[i].alfa));
    if (dist > 0.5) split.push_back(i + 1);
  }
}

/* LineConversion, lineFitting.cpp:586-638 */
void conversion(std::vector<Line>& lines) {
  std::vector<Line> keep;
  for (size_t i = 0; i < lines.size(); i++) {
    const Line& l = lines[i];
    if (l.C[0] < 0 || l.C[3] < 0) continue;
    if (std::isnan(l.C[0]) || std::isnan(l.C[1]) || std::isnan(l.C[2]) || std::isnan(l.C[3])) continue;
    if (l.alfa == 0 && l.r == 0) continue;
    if (l.C[0] > 0.01) continue;
    keep.push_back(l);
  }
  for (size_t i = 0; i < keep.size(); i++) canonical(keep[i].alfa, keep[i].r);
  lines.swap(keep);
}

std::vector<Line> extract(std::vector<PolarPoint>& pts) {
  /* qsort by alfa + PI (lineFitting.cpp:520-539, 642); ties keep input order here */
  std::stable_sort(pts.begin(), pts.end(), [](const PolarPoint& a, const PolarPoint& b) { return (a.alfa + PI) < (b.alfa + PI); });
  std::vector<int> split;
  segmentation(pts, split);
  std::vector<Line> lines;
  if (!split.empty()) {
    std::rotate(pts.begin(), pts.begin() + split.back(), pts.end());
    split.clear();
    segmentation(pts, split);
  }
  if (!split.empty()) {
    const int ns = (int)split.size() + 1;
    for (int s = 0; s < ns; s++) {
      const int lo = (s == 0) ? 0 : split[s - 1];
      const int hi = (s == ns - 1) ? (int)pts.size() : split[s];
      rdp(pts.data() + lo, hi - lo, lines);
    }
  } else {
    rdp(pts.data(), (int)pts.size(), lines);
  }
  conversion(lines);
  return lines;
}

}  // namespace

extern "C" {

/* Same contract as ref_extract_lines (oracle/ref_lines_harness.cpp): data = n_pairs x (r, angle) float32;
 * out = 10 doubles per line: alfa, r, C_AR[4], interval0 (alfa, r), interval1 (alfa, r). */
int lxo_extract(int n_pairs, const float* data, int max_lines, double* out) {
  std::vector<PolarPoint> pts;
  for (int i = 0; i < 2 * n_pairs; i += 2) {                       /* main.cpp:46-62 */
    if (data[i] > 0.05) {
      PolarPoint q;
      q.alfa = data[i + 1] - M_PI;
      q.r = data[i];
      q.weight = 1;
      q.variance = 0.01;
      pts.push_back(q);
    }
  }
  std::vector<Line> lines;
  if (pts.size() >= 2) lines = extract(pts);
  const int n = (int)lines.size();
  for (int i = 0; i < n && i < max_lines; i++) {
    double a = lines[i].alfa + M_PI;                                /* main.cpp:66-69 */
    a = a > M_PI ? a - 2.0 * M_PI : a;
    double* o = out + 10 * i;
    o[0] = a; o[1] = lines[i].r;
    for (int k = 0; k < 4; k++) o[2 + k] = lines[i].C[k];
    for (int k = 0; k < 2; k++) { o[6 + 2 * k] = lines[i].ia[k]; o[7 + 2 * k] = lines[i].ir[k]; }
  }
  return n;
}

/* one fit (lineFitting.cpp:267-304) on explicit polar points with unit weights: out = (alfa, r) */
void lxo_fit(int n, const double* alfa, const double* r, double* out) {
  std::vector<PolarPoint> p(n);
  for (int i = 0; i < n; i++) { p[i].alfa = alfa[i]; p[i].r = r[i]; p[i].weight = 1; p[i].variance = 0.01; }
  Line l = fit(p.data(), n);
  out[0] = l.alfa; out[1] = l.r;
}

}  /* extern "C" */
