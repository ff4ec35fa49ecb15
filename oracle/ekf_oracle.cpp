/* ekf_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  Nothing in slam_ros_b200/ links, loads or calls this.
 *
 * CPU restatement ("structured oracle") of the reference EKF-SLAM hot path,
 * slam_ros/Robot.cpp:126-943 (Robot::localize) with runtime capacity and heap
 * storage, so that it can run at 1k / 10k / 40k landmarks where the literal
 * reference cannot (Robot.h:13-14 fixes LINESIZE=100; every temporary is an
 * n x n stack array; prediction is a dense n^3 dgemm).
 *
 * It follows the reference operation for operation: every floating-point
 * expression below is the scalar expansion of the GSL reference-BLAS loop the
 * reference calls at the cited line (k-outer NN/TN with zero skip, dot-product
 * NT -- see oracle/gsl_shim/gsl_shim.h), restricted to the entries that are
 * structurally non-zero.  FULL n x n, NON-symmetrised covariance, exactly like
 * the reference.  Build with -ffp-contract=off.
 *
 * Pinning: tests/test_oracle_vs_literal.py checks this file BITWISE against the
 * literal reference (oracle/_ref/libslamref.so = unmodified Robot.cpp compiled
 * over the GSL shim) at LINESIZE=100 over 1k-step scenarios including
 * augmentation and the map reset.  The reference's only deviation: the one-token
 * fix of the R[i] typo at Robot.cpp:303 -- R is the observed
 * line's 2x2 C_AR, as at Robot.cpp:807-810.
 *
 * Parity status: the reference holds no tests or golden vectors for this path
 * and GSL itself is absent from this image => "parity pinned to the reference
 * SOURCE compiled over a restated GSL", not to a libgsl binary.
 *
 * OpenMP is used only for element-wise independent loops (row sweeps); results
 * are bitwise independent of the thread count.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Oracle {
  int cap;            /* LINESIZE */
  int n;              /* SLAMSIZE = 3 + 2*cap */
  double gate;        /* MAHALANOBIS, Robot.h:15 */
  double enc_noise;   /* ENCODERNOISE, Robot.h:17 */
  int headroom;       /* the "10" of Robot.cpp:893 */
  double* y;          /* Robot.h:26 */
  double* P;          /* Robot.h:62, row-major n x n (this is P_pre and P_t0 at once) */
  int L;              /* savedLineCount */
  double pose[3];     /* xPos, yPos, thetaPos */
  double x_pre[3];
  /* per-scan */
  std::vector<int> matched;     /* matchSavedIndexes, Robot.cpp:294 */
  std::vector<int> extra;       /* indices (into the scan) of extraLines, Robot.cpp:291 */
  int matches;                  /* matchesNum */
  /* statistics */
  double min_margin;            /* min |sqrt|d2| - gate| over every gate evaluated */
  long long gates, total_matches, resets;
  int threads;
  /* scratch */
  std::vector<double> K, KS;
  /* last gate evaluated (diagnostic tap) */
  double last_S[4], last_Sinv[4], last_v[2], last_d2;
};

/* Robot.cpp:62-71, verbatim semantics (Q6: wrong for |rad| > 2*pi, reproduced) */
inline void normalize_radian(double& rad) {
  if (rad > M_PI) {
    rad = rad - (2.0 * M_PI + std::floor(rad / (2.0 * M_PI)) * 2.0 * M_PI);
  } else if (rad < -M_PI) {
    rad = rad + (2.0 * M_PI + std::floor(std::abs(rad) / (2.0 * M_PI)) * 2.0 * M_PI);
  }
}

/* acc += t*b with the reference-BLAS "skip when t == 0" rule of NN/TN dgemm */
inline void axpy_skip(double& acc, double t, double b) { if (t != 0.0) acc += t * b; }

/* gsl_linalg_LU_decomp + LU_invert on a 2x2 (Robot.cpp:431-457).  Returns 0 or GSL_EDOM(1). */
int inv2x2_lu(const double S[4], double Si[4]) {
  double a[4] = {S[0], S[1], S[2], S[3]};
  int p0 = 0, p1 = 1;
  if (std::fabs(a[2]) > std::fabs(a[0])) {           /* first-largest pivot in column 0 */
    std::swap(a[0], a[2]); std::swap(a[1], a[3]); p0 = 1; p1 = 0;
  }
  if (a[0] != 0.0) {
    const double l = a[2] / a[0];
    a[2] = l;
    a[3] = a[3] - l * a[1];
  }
  if (a[0] == 0.0 || a[3] == 0.0) { Si[0] = Si[1] = Si[2] = Si[3] = 0.0; return 1; }   /* S_inv stays {0,0,0,0} (Robot.cpp:443) */
  for (int col = 0; col < 2; ++col) {
    const double e[2] = {col == 0 ? 1.0 : 0.0, col == 1 ? 1.0 : 0.0};
    double x0 = e[p0], x1 = e[p1];
    { double s = x1; s -= a[2] * x0; x1 = s; }       /* unit-lower forward substitution */
    x1 = x1 / a[3];                                  /* upper back substitution */
    { double s = x0; s -= a[1] * x1; x0 = s / a[0]; }
    Si[0 * 2 + col] = x0; Si[1 * 2 + col] = x1;
  }
  return 0;
}

struct Gate {
  double c, s, g;       /* cos(alpha_j), sin(alpha_j), x*s - y*c  (Robot.cpp:373-379) */
  double S[4], Si[4];
  double v[2];          /* innovation after the wraps */
  double d2;
  int singular;
};

/* Robot.cpp:367-489 for one (line, landmark j) pair against the CURRENT P / x_pre / y. */
void eval_gate(const Oracle* o, int j, const double z_in[2], const double R[4], Gate* G) {
  const int n = o->n, a = 3 + 2 * j, b = a + 1;
  const double* P = o->P;
  const double yj = o->y[a];
  const double c = std::cos(yj), s = std::sin(yj);
  const double H10 = -c, H11 = -s;
  const double g = o->x_pre[0] * s - o->x_pre[1] * c;
  G->c = c; G->s = s; G->g = g;
  const int cols[5] = {0, 1, 2, a, b};
  double HP0[5], HP1[5];
  for (int t = 0; t < 5; ++t) {                       /* Robot.cpp:397  H(2xn) * P  (NN) */
    const int col = cols[t];
    double h0 = 0.0, h1 = 0.0;
    /* k = 0 */ axpy_skip(h1, H10, P[(size_t)0 * n + col]);
    /* k = 1 */ axpy_skip(h1, H11, P[(size_t)1 * n + col]);
    /* k = 2 */ axpy_skip(h0, -1.0, P[(size_t)2 * n + col]);
    /* k = a */ axpy_skip(h0, 1.0, P[(size_t)a * n + col]); axpy_skip(h1, g, P[(size_t)a * n + col]);
    /* k = b */ axpy_skip(h1, 1.0, P[(size_t)b * n + col]);
    HP0[t] = h0; HP1[t] = h1;
  }
  const double* HP[2] = {HP0, HP1};
  for (int p = 0; p < 2; ++p) {                       /* Robot.cpp:401  HP * H'  (NT) */
    double t0 = 0.0;
    t0 += HP[p][2] * -1.0;
    t0 += HP[p][3] * 1.0;
    double t1 = 0.0;
    t1 += HP[p][0] * H10;
    t1 += HP[p][1] * H11;
    t1 += HP[p][3] * g;
    t1 += HP[p][4] * 1.0;
    G->S[p * 2 + 0] = 0.0 + t0;
    G->S[p * 2 + 1] = 0.0 + t1;
  }
  for (int t = 0; t < 4; ++t) G->S[t] += R[t];        /* Robot.cpp:405 */
  const double m0 = o->y[a], m1 = o->y[b];            /* Robot.cpp:417-426 */
  double h[2] = {m0 - o->x_pre[2], m1 - (o->x_pre[0] * std::cos(m0) + o->x_pre[1] * std::sin(m0))};
  normalize_radian(h[0]);
  G->singular = inv2x2_lu(G->S, G->Si);
  double z[2] = {z_in[0], z_in[1]};
  z[0] -= h[0]; z[1] -= h[1];                         /* Robot.cpp:465 */
  if (std::abs(z[0] - 2.0 * M_PI) < std::abs(z[0])) z[0] -= 2.0 * M_PI;       /* Robot.cpp:471-475 */
  else if (std::abs(z[0] + 2.0 * M_PI) < std::abs(z[0])) z[0] += 2.0 * M_PI;
  G->v[0] = z[0]; G->v[1] = z[1];
  double w[2] = {0.0, 0.0};                           /* Robot.cpp:479  z' * Sinv  (TN) */
  for (int k = 0; k < 2; ++k) { axpy_skip(w[0], z[k], G->Si[k * 2 + 0]); axpy_skip(w[1], z[k], G->Si[k * 2 + 1]); }
  double d2 = 0.0;                                    /* Robot.cpp:483  (NN) */
  axpy_skip(d2, w[0], z[0]); axpy_skip(d2, w[1], z[1]);
  G->d2 = d2;
}

bool gate_rejects(const Oracle* o, const Gate& G) {  /* Robot.cpp:489 */
  return std::sqrt(std::abs(G.d2)) > o->gate;
}

/* Robot.cpp:516-602: gain, covariance update, state update for matched landmark j. */
void apply_update(Oracle* o, int j, const Gate& G) {
  const int n = o->n, nl = 3 + 2 * o->L, a = 3 + 2 * j, b = a + 1;
  double* P = o->P;
  double* K = o->K.data();
  double* KS = o->KS.data();
  const double H10 = -G.c, H11 = -G.s, g = G.g;
  const double* Si = G.Si; const double* S = G.S;
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int r = 0; r < nl; ++r) {
    const double* Pr = P + (size_t)r * n;
    double t0 = 0.0;                                  /* Robot.cpp:522  P * H'  (NT) */
    t0 += Pr[2] * -1.0;
    t0 += Pr[a] * 1.0;
    double t1 = 0.0;
    t1 += Pr[0] * H10;
    t1 += Pr[1] * H11;
    t1 += Pr[a] * g;
    t1 += Pr[b] * 1.0;
    const double ph0 = 0.0 + t0, ph1 = 0.0 + t1;
    double k0 = 0.0, k1 = 0.0;                        /* Robot.cpp:526  PHt * Sinv  (NN) */
    axpy_skip(k0, ph0, Si[0]); axpy_skip(k1, ph0, Si[1]);
    axpy_skip(k0, ph1, Si[2]); axpy_skip(k1, ph1, Si[3]);
    K[2 * r] = k0; K[2 * r + 1] = k1;
    double s0 = 0.0, s1 = 0.0;                        /* Robot.cpp:560  K * S  (NN) */
    axpy_skip(s0, k0, S[0]); axpy_skip(s1, k0, S[1]);
    axpy_skip(s0, k1, S[2]); axpy_skip(s1, k1, S[3]);
    KS[2 * r] = s0; KS[2 * r + 1] = s1;
  }
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int r = 0; r < nl; ++r) {                      /* Robot.cpp:564-568  P -= KS * K'  (NT, full matrix) */
    double* Pr = P + (size_t)r * n;
    const double ks0 = KS[2 * r], ks1 = KS[2 * r + 1];
    for (int q = 0; q < nl; ++q) {
      double t = 0.0;
      t += ks0 * K[2 * q];
      t += ks1 * K[2 * q + 1];
      Pr[q] -= (0.0 + t);
    }
  }
  o->y[0] = o->x_pre[0]; o->y[1] = o->x_pre[1]; o->y[2] = o->x_pre[2];   /* Robot.cpp:579-589 */
  for (int r = 0; r < nl; ++r) {
    double t = 0.0;
    axpy_skip(t, K[2 * r], G.v[0]);
    axpy_skip(t, K[2 * r + 1], G.v[1]);
    o->y[r] += t;
  }
  normalize_radian(o->y[2]);                          /* Robot.cpp:596-602 */
  o->pose[0] = o->y[0]; o->pose[1] = o->y[1]; o->pose[2] = o->y[2];
  o->x_pre[0] = o->y[0]; o->x_pre[1] = o->y[1]; o->x_pre[2] = o->y[2];
}

/* Robot.cpp:148-258 in structured form. */
void predict(Oracle* o, const double u[3]) {
  const int n = o->n, nl = 3 + 2 * o->L;
  double* P = o->P;
  const double* x = o->pose;
  const double ang = x[2] + u[2] / 2.0;
  const double ca = std::cos(ang), sa = std::sin(ang);
  o->x_pre[0] = x[0] + u[0] * ca;                     /* Robot.cpp:148 */
  o->x_pre[1] = x[1] + u[0] * sa;
  o->x_pre[2] = x[2] + u[2];
  const double F02 = -u[0] * sa, F12 = u[0] * ca;     /* Robot.cpp:157,160 */
  /* Robot.cpp:242  T = Fx * P (NN): only rows 0..2 are not plain copies */
  for (int j = 0; j < nl; ++j) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    axpy_skip(t0, 1.0, P[(size_t)0 * n + j]);
    axpy_skip(t1, 1.0, P[(size_t)1 * n + j]);
    axpy_skip(t0, F02, P[(size_t)2 * n + j]);
    axpy_skip(t1, F12, P[(size_t)2 * n + j]);
    axpy_skip(t2, 1.0, P[(size_t)2 * n + j]);
    P[(size_t)0 * n + j] = t0; P[(size_t)1 * n + j] = t1; P[(size_t)2 * n + j] = t2;
  }
  /* Robot.cpp:246  P_pre = T * Fx' (NT): only columns 0,1 are not plain copies */
  for (int i = 0; i < nl; ++i) {
    double* Ti = P + (size_t)i * n;
    double c0 = 0.0; c0 += Ti[0] * 1.0; c0 += Ti[2] * F02;
    double c1 = 0.0; c1 += Ti[1] * 1.0; c1 += Ti[2] * F12;
    Ti[0] = 0.0 + c0; Ti[1] = 0.0 + c1;
  }
  /* Robot.cpp:180-188 Fu, :215-218 Q, :250-258 P_pre += Fu*Q*Fu' */
  const double Fu[9] = {ca, 0.0, -u[0] * sa / 2.0,
                        sa, 1.0, u[0] * ca / 2.0,
                        0.0, 0.0, 1.0};
  const double Q[9] = {o->enc_noise * (-1.0 / (1 + std::abs(u[0])) + 1), 0, 0,
                       0, 2 * o->enc_noise * (-1.0 / (1 + std::abs(u[0])) + 1), 0,
                       0, 0, o->enc_noise * (-1.0 / (1 + std::abs(u[0])) + 1)};
  double FQ[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 3; ++i) {
      const double t = 1.0 * Fu[i * 3 + k];
      if (t != 0.0) for (int j = 0; j < 3; ++j) FQ[i * 3 + j] += t * Q[k * 3 + j];
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double t = 0.0;
      for (int k = 0; k < 3; ++k) t += FQ[i * 3 + k] * Fu[j * 3 + k];
      P[(size_t)i * n + j] += (0.0 + 1.0 * t);
    }
}

void begin_scan(Oracle* o) { o->matched.clear(); o->extra.clear(); o->matches = 0; }

/* Robot.cpp:298-645 for one observed line: returns matched landmark index or -1 (line becomes "extra"). */
int process_line(Oracle* o, int line_idx, const double z[2], const double R[4]) {
  if (o->L == 0) { o->extra.push_back(line_idx); return -1; }      /* Robot.cpp:308-310 */
  for (int j = 0; j < o->L; ++j) {
    if (std::find(o->matched.begin(), o->matched.end(), j) != o->matched.end()) {   /* :315-330 */
      if (j == o->L - 1) { o->extra.push_back(line_idx); return -1; }
      continue;
    }
    Gate G;
    eval_gate(o, j, z, R, &G);
    o->gates++;
    const double d = std::sqrt(std::abs(G.d2));
    const double margin = std::fabs(d - o->gate);
    if (margin < o->min_margin) o->min_margin = margin;
    if (gate_rejects(o, G)) {                                         /* :489-498 */
      if (j == o->L - 1) { o->extra.push_back(line_idx); return -1; }
      continue;
    }
    o->matched.push_back(j);                                          /* :501-504 */
    o->matches++; o->total_matches++;
    std::memcpy(o->last_S, G.S, sizeof G.S); std::memcpy(o->last_Sinv, G.Si, sizeof G.Si);
    o->last_v[0] = G.v[0]; o->last_v[1] = G.v[1]; o->last_d2 = G.d2;
    apply_update(o, j, G);
    return j;
  }
  return -1; /* unreachable: the last j always returns */
}

/* Robot.cpp:792-866 for one unmatched line (alfa, r robot frame; R = its C_AR). */
void add_line(Oracle* o, double alfa, double r, const double R[4]) {
  const int n = o->n, L = o->L, l = 3 + 2 * L;
  double* P = o->P;
  r += (o->pose[0] * std::cos(alfa) + o->pose[1] * std::sin(alfa));  /* :792 (robot-frame angle, Q8) */
  alfa += o->pose[2];                                                 /* :793 */
  const double cw = std::cos(alfa), sw = std::sin(alfa);
  const double Gx[6] = {0, 0, 1, cw, sw, 0};
  const double Gl[4] = {1.0, 0, o->y[1] * cw - o->y[0] * sw, 1};
  normalize_radian(alfa);                                             /* :801 */
  o->y[l] = alfa; o->y[l + 1] = r;
  double GP[6] = {0, 0, 0, 0, 0, 0};                                  /* :823 Gx * Prr (NN) */
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 2; ++i) {
      const double t = 1.0 * Gx[i * 3 + k];
      if (t != 0.0) for (int j = 0; j < 3; ++j) GP[i * 3 + j] += t * P[(size_t)k * n + j];
    }
  double GPG[4];                                                      /* :827 (NT) */
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      double t = 0.0;
      for (int k = 0; k < 3; ++k) t += GP[i * 3 + k] * Gx[j * 3 + k];
      GPG[i * 2 + j] = 0.0 + 1.0 * t;
    }
  double GR[4] = {0, 0, 0, 0};                                        /* :831 Gl * R (NN) */
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < 2; ++i) {
      const double t = 1.0 * Gl[i * 2 + k];
      if (t != 0.0) for (int j = 0; j < 2; ++j) GR[i * 2 + j] += t * R[k * 2 + j];
    }
  double GRG[4];                                                      /* :835 (NT) */
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      double t = 0.0;
      for (int k = 0; k < 2; ++k) t += GR[i * 2 + k] * Gl[j * 2 + k];
      GRG[i * 2 + j] = 0.0 + 1.0 * t;
    }
  for (int t = 0; t < 4; ++t) GPG[t] += GRG[t];                       /* :839 */
  P[(size_t)l * n + l] = GPG[0]; P[(size_t)l * n + l + 1] = GPG[1];   /* :844-845 */
  P[(size_t)(l + 1) * n + l] = GPG[2]; P[(size_t)(l + 1) * n + l + 1] = GPG[3];
  for (int j = 0; j < l; ++j) {                                       /* :856 Gx * [Prr Prm] (NN, in place) */
    double r0 = 0.0, r1 = 0.0;
    for (int k = 0; k < 3; ++k) {
      axpy_skip(r0, 1.0 * Gx[0 * 3 + k], P[(size_t)k * n + j]);
      axpy_skip(r1, 1.0 * Gx[1 * 3 + k], P[(size_t)k * n + j]);
    }
    P[(size_t)l * n + j] = r0; P[(size_t)(l + 1) * n + j] = r1;
  }
  for (int j = 0; j < l; ++j) {                                       /* :860 transpose into the column block */
    P[(size_t)j * n + l] = P[(size_t)l * n + j];
    P[(size_t)j * n + l + 1] = P[(size_t)(l + 1) * n + j];
  }
  o->L = L + 1;                                                       /* :866 */
}

/* Robot.cpp:702-716, :776-866 (augmentation driver), :893-904 (reset).  z/R are the scan's lines. */
int end_scan(Oracle* o, int m, const double* z, const double* R) {
  int status = 0;
  if (m == 0 || o->matches == 0) {
    o->y[0] = o->x_pre[0]; o->y[1] = o->x_pre[1]; o->y[2] = o->x_pre[2];
    o->pose[0] = o->y[0]; o->pose[1] = o->y[1]; o->pose[2] = o->y[2];
    normalize_radian(o->pose[2]);
  }
  for (size_t e = 0; e < o->extra.size(); ++e) {
    const int i = o->extra[e];
    if (o->L >= o->cap) { status = 2; break; }     /* the reference overruns y[] here (Q4); defined as: drop + flag */
    add_line(o, z[2 * i], z[2 * i + 1], R + 4 * i);
  }
  if (o->L > o->cap - o->headroom) {               /* Robot.cpp:893-904 */
    const int n = o->n, nl = 3 + 2 * o->L;
    o->L = 0;
    for (int i = 3; i < nl; ++i) o->y[i] = 0;
    for (int i = 0; i < nl; ++i)
      for (int j = 0; j < nl; ++j)
        if (i >= 3 || j >= 3) o->P[(size_t)i * n + j] = 0;
    o->resets++;
  }
  return status;
}

}  // namespace

extern "C" {

void* ekfo_create(int capacity_lines, double gate, double encoder_noise, int reset_headroom) {
  Oracle* o = new Oracle();
  o->cap = capacity_lines; o->n = 3 + 2 * capacity_lines;
  o->gate = gate; o->enc_noise = encoder_noise; o->headroom = reset_headroom;
  o->y = (double*)std::calloc((size_t)o->n, sizeof(double));
  o->P = (double*)std::calloc((size_t)o->n * o->n, sizeof(double));
  if (!o->y || !o->P) { std::free(o->y); std::free(o->P); delete o; return 0; }
  o->L = 0;
  o->pose[0] = o->pose[1] = o->pose[2] = 0.0;
  o->x_pre[0] = o->x_pre[1] = o->x_pre[2] = 0.0;
  o->P[0] = 0.05; o->P[(size_t)o->n + 1] = 0.05; o->P[(size_t)2 * o->n + 2] = 0;   /* Robot.cpp:27-30 */
  o->matches = 0; o->min_margin = HUGE_VAL; o->gates = 0; o->total_matches = 0; o->resets = 0;
  o->threads = 1;
  o->K.assign((size_t)2 * o->n, 0.0); o->KS.assign((size_t)2 * o->n, 0.0);
  return o;
}
void ekfo_destroy(void* h) { Oracle* o = (Oracle*)h; if (!o) return; std::free(o->y); std::free(o->P); delete o; }
void ekfo_set_threads(void* h, int t) {
  Oracle* o = (Oracle*)h;
#ifdef _OPENMP
  o->threads = t > 0 ? t : omp_get_max_threads();
#else
  (void)t; o->threads = 1;
#endif
}
int ekfo_get_threads(void* h) { return ((Oracle*)h)->threads; }
/* adopt a state prepared elsewhere (bench.py: the -O0 build continues from the -O2 build's seeded map); y and P are
 * written through ekfo_y_ptr / ekfo_P_ptr */
void ekfo_set_lines(void* h, int L) { Oracle* o = (Oracle*)h; if (L >= 0 && L <= o->cap) o->L = L; }
void ekfo_set_pose(void* h, const double pose[3]) { Oracle* o = (Oracle*)h; std::memcpy(o->pose, pose, sizeof o->pose); }

/* primitives mirroring include/ekf.h */
void ekfo_predict(void* h, const double u[3], double x_pre[3]) {
  Oracle* o = (Oracle*)h; begin_scan(o); predict(o, u);
  if (x_pre) std::memcpy(x_pre, o->x_pre, sizeof o->x_pre);
}
/* first-fit association WITHOUT applying the update; innov = innovation of the winner */
int ekfo_associate(void* h, const double z[2], const double R[4], double innov[2], double* d2) {
  Oracle* o = (Oracle*)h;
  for (int j = 0; j < o->L; ++j) {
    if (std::find(o->matched.begin(), o->matched.end(), j) != o->matched.end()) continue;
    Gate G; eval_gate(o, j, z, R, &G);
    if (!gate_rejects(o, G)) { if (innov) { innov[0] = G.v[0]; innov[1] = G.v[1]; } if (d2) *d2 = G.d2; return j; }
  }
  return -1;
}
/* gate of one explicit pair (diagnostic tap for kernel unit tests) */
void ekfo_gate_pair(void* h, int j, const double z[2], const double R[4], double S[4], double Sinv[4], double v[2], double* d2) {
  Oracle* o = (Oracle*)h; Gate G; eval_gate(o, j, z, R, &G);
  std::memcpy(S, G.S, sizeof G.S); std::memcpy(Sinv, G.Si, sizeof G.Si); v[0] = G.v[0]; v[1] = G.v[1]; *d2 = G.d2;
}
void ekfo_update(void* h, int j, const double z[2], const double R[4]) {
  Oracle* o = (Oracle*)h; Gate G; eval_gate(o, j, z, R, &G);
  o->matched.push_back(j); o->matches++; o->total_matches++;
  apply_update(o, j, G);
}
/* extraLines.push_back(lines[line_idx]) and the end-of-scan block, for tests that drive the primitives */
void ekfo_queue(void* h, int line_idx) { ((Oracle*)h)->extra.push_back(line_idx); }
int ekfo_end(void* h, int m, const double* z, const double* R) { return end_scan((Oracle*)h, m, z, R); }
void ekfo_last_gain(void* h, double* K, double* KS) {   /* n x 2 each, of the most recent update */
  Oracle* o = (Oracle*)h;
  std::memcpy(K, o->K.data(), sizeof(double) * 2 * o->n); std::memcpy(KS, o->KS.data(), sizeof(double) * 2 * o->n);
}

/* one full Robot::localize with the odometry given directly as u (u[1] unused, as in the reference) */
int ekfo_scan(void* h, const double u[3], int m, const double* z, const double* R, int* j_out) {
  Oracle* o = (Oracle*)h;
  begin_scan(o);
  predict(o, u);
  for (int i = 0; i < m; ++i) {
    const int j = process_line(o, i, z + 2 * i, R + 4 * i);
    if (j_out) j_out[i] = j;
  }
  return end_scan(o, m, z, R);
}
/* the literal entry: odometry from (pose - encoder), Robot.cpp:140-145 (Q5) */
int ekfo_localize(void* h, int m, const double* z, const double* R, const double encoder[3], int* j_out) {
  Oracle* o = (Oracle*)h;
  double u[3] = {0, 0, 0};
  u[2] = o->pose[2] - encoder[2];
  const double dX = o->pose[0] - encoder[0], dY = o->pose[1] - encoder[1];
  u[0] = std::sqrt(dX * dX + dY * dY);
  return ekfo_scan(h, u, m, z, R, j_out);
}

int ekfo_n(void* h) { return ((Oracle*)h)->n; }
int ekfo_lines(void* h) { return ((Oracle*)h)->L; }
double* ekfo_y_ptr(void* h) { return ((Oracle*)h)->y; }
double* ekfo_P_ptr(void* h) { return ((Oracle*)h)->P; }
void ekfo_get_pose(void* h, double pose[3]) { std::memcpy(pose, ((Oracle*)h)->pose, 3 * sizeof(double)); }
void ekfo_get_xpre(void* h, double x[3]) { std::memcpy(x, ((Oracle*)h)->x_pre, 3 * sizeof(double)); }
void ekfo_stats(void* h, double* min_margin, long long* gates, long long* matches, long long* resets) {
  Oracle* o = (Oracle*)h;
  if (min_margin) *min_margin = o->min_margin;
  if (gates) *gates = o->gates;
  if (matches) *matches = o->total_matches;
  if (resets) *resets = o->resets;
}
/* trace, sum and sum of squares of the live covariance as libekfcuda reports it (ekf_cov_stats: upper triangle
 * mirrored), for full-size checks where the matrix itself (51 GB at 40k landmarks) cannot be compared or copied */
void ekfo_upper_stats(void* h, double* trace, double* sum, double* sumsq) {
  Oracle* o = (Oracle*)h; const int nl = 3 + 2 * o->L; const size_t n = o->n;
  double tr = 0.0, sm = 0.0, sq = 0.0;
#pragma omp parallel for reduction(+ : tr, sm, sq) schedule(dynamic, 64) num_threads(o->threads)
  for (int r = 0; r < nl; ++r) {
    const double* Pr = o->P + (size_t)r * n;
    double s = 0.0, q = 0.0;
    for (int c = r + 1; c < nl; ++c) { s += Pr[c]; q += Pr[c] * Pr[c]; }
    tr += Pr[r]; sm += 2.0 * s + Pr[r]; sq += 2.0 * q + Pr[r] * Pr[r];
  }
  if (trace) *trace = tr; if (sum) *sum = sm; if (sumsq) *sumsq = sq;
}
/* copy out the live (nl x nl) corner, row-major with leading dimension nl */
void ekfo_get_live(void* h, double* y, double* P) {
  Oracle* o = (Oracle*)h; const int nl = 3 + 2 * o->L, n = o->n;
  if (y) std::memcpy(y, o->y, sizeof(double) * nl);
  if (P) for (int i = 0; i < nl; ++i) std::memcpy(P + (size_t)i * nl, o->P + (size_t)i * n, sizeof(double) * nl);
}

/* Robot::getEllipse (Robot.cpp:73-124) in closed form on P[0:2,0:2]; axes = 2*sqrt(5.991*|lambda|),
 * sorted by |lambda| ascending; angle = atan2(v_x, v_y) of the LAST (largest) eigenvector.  Float outputs. */
int ekfo_get_ellipse(void* h, float axii[2], float* angle) {
  Oracle* o = (Oracle*)h; const int n = o->n;
  const double a = o->P[0], b = o->P[1], c = o->P[n], d = o->P[n + 1];
  const double tr = a + d, half = 0.5 * (a - d), disc = half * half + b * c;
  if (disc < 0.0) return 0;
  const double rt = std::sqrt(disc);
  double l[2] = {0.5 * tr + rt, 0.5 * tr - rt};
  double v[2][2];
  for (int k = 0; k < 2; ++k) {
    double vx, vy;
    if (b != 0.0) { vx = b; vy = l[k] - a; }
    else if (c != 0.0) { vx = l[k] - d; vy = c; }
    else { vx = (k == 0) == (a >= d) ? 1.0 : 0.0; vy = 1.0 - vx; }
    const double nrm = std::sqrt(vx * vx + vy * vy);
    if (nrm > 0.0) { vx /= nrm; vy /= nrm; }
    v[k][0] = vx; v[k][1] = vy;
  }
  if (std::fabs(l[1]) < std::fabs(l[0])) { std::swap(l[0], l[1]); std::swap(v[0][0], v[1][0]); std::swap(v[0][1], v[1][1]); }
  for (int k = 0; k < 2; ++k) axii[k] = 2.f * std::sqrt(5.991 * std::abs(l[k]));
  *angle = std::atan2(v[1][0], v[1][1]);
  return 1;
}

}  /* extern "C" */
