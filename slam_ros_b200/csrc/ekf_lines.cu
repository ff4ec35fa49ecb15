/* ekf_lines.cu -- line extraction on the device: what the node does with one
 * `mappingPoints` payload before Robot::localize -- slam_ros/main.cpp:37-71 (`mapping_cb`, SIMULATIONOFF branch)
 * and LineExtraction (slam_ros/lineFitting.cpp:640-702: sort, segmentation :541-584, the recursive split
 * simplifyPath.cpp:108-177, the fit lineFitting.cpp:267-304, its finite-difference covariance :379-450, the
 * end points simplifyPath.cpp:59-105 + lineFitting.cpp:44-51, LineConversion :586-638).
 *
 * Same algorithm and decisions, re-organised for a GPU:
 *   - the fit's O(p^2) pair sums  sum_{i<j} r_i r_j sin/cos(a_i + a_j)  are evaluated in their closed form over the
 *     Cartesian points X = r cos a, Y = r sin a:  N = (2/p) sum_1 + (1/p) sum_2 = -2 S_xy,  D = -(S_xx - S_yy)
 *     (centred second moments, two passes), and  sum_i r_i cos(a_i - alfa) = cos(alfa) SX + sin(alfa) SY.
 *     alfa = 0.5 atan2(N, D) exactly as at lineFitting.cpp:297; the centred form is, if anything, better
 *     conditioned than the reference's own summation, so (alfa, r) agree with it to ~1e-13;
 *   - the covariance keeps the reference's forward difference with eps = 1e-6 on every range (one refit per
 *     point, O(p) each); the refits on the ANGLES multiply `1/12*1.5` == 0 (integer division, :419) and are
 *     not evaluated.  The reference's own finite differences carry ~1e-7 relative rounding noise, so C_AR is
 *     compared at 1e-4 relative;
 *   - segments are independent: one thread block per segment walks its split recursion with an explicit stack
 *     (block-wide reductions per node); leaves are keyed by their first point, which IS the reference's output
 *     order (segment order, then left-to-right), so a final single-block pass filters and compacts them.
 * Indeterminate reads of the reference are fixed as in oracle/lines_oracle.cpp (zero-initialised accumulators,
 * diagonal C_x).  PI below is the truncated constant of lineFitting.h:12; main.cpp uses M_PI.
 */
#include "../../include/ekf.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#define LX_MAX_POINTS 4096            /* returns per payload (a 0.09-degree scanner) */
#define LX_PREP_THREADS 1024
#define LX_ITEMS (LX_MAX_POINTS / LX_PREP_THREADS)
#define LX_SEG_THREADS 128
#define LX_PI 3.14159265
#define LX_MPI 3.14159265358979323846

namespace {

struct LxLeaf {            /* keyed by the index of the leaf's first point */
  double alfa, r, c0, c3, ia0, ir0, ia1, ir1;
  int valid, pad;
};

struct LxBuffers {
  const float* data;       /* payload: n_pairs x (r, angle) */
  double *a, *r, *X, *Y, *ca, *sa;   /* sorted (and rotated) points: angle, range, r cos a, r sin a, cos a, sin a */
  int* seg;                /* seg[0] = number of segments, seg[1 + s] = first point of segment s, seg[1 + nseg] = n points */
  LxLeaf* leaf;            /* [LX_MAX_POINTS] */
  double* out;             /* [max_lines][10] */
  double* z;               /* [max_lines][2]  (alfa, r) for ekf_scan_device */
  double* R;               /* [max_lines][4]  C_AR */
  int* count;              /* lines found */
};

/* ---- block-wide helpers (fixed reduction order: deterministic) ---- */
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* sbuf) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sbuf[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < THREADS / 32; ++w) t += sbuf[w];
  return t;
}
/* N sums in ONE pass (one pair of barriers instead of N): v[] in, totals out in v[]; sbuf holds N * THREADS/32 doubles */
template <int THREADS, int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double* sbuf) {
#pragma unroll
  for (int k = 0; k < N; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) sbuf[k * (THREADS / 32) + (threadIdx.x >> 5)] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double t = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) t += sbuf[k * (THREADS / 32) + w];
    v[k] = t;
  }
}
template <int THREADS>
__device__ __forceinline__ int block_sum_int(int v, int* sbuf) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sbuf[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < THREADS / 32; ++w) t += sbuf[w];
  return t;
}
/* exclusive prefix sum of 0/1 flags over the block; returns this thread's offset, *total the sum */
template <int THREADS>
__device__ __forceinline__ int block_scan_flags(int flag, int* sbuf, int* total) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sbuf[w] = __popc(m);
  __syncthreads();
  int base = 0, tot = 0;
  for (int k = 0; k < THREADS / 32; ++k) { const int c = sbuf[k]; if (k < w) base += c; tot += c; }
  *total = tot;
  return base + __popc(m & ((1u << lane) - 1u));
}

__host__ __device__ __forceinline__ int lx_npad(int n_pairs) {
  int npad = LX_PREP_THREADS;
  while (npad < n_pairs && npad < LX_MAX_POINTS) npad <<= 1;
  return npad;
}

/* ------------------------------------------------------------------------------------------------ */
/* main.cpp:46-62 (polar points), lineFitting.cpp:642 (sort by alfa + PI), :541-584 (segmentation),
 * :645-649 (rotate so that the scan starts at a break, segment again). */
__global__ void __launch_bounds__(LX_PREP_THREADS) k_lx_prepare(LxBuffers b, int n_pairs) {
  extern __shared__ unsigned char lx_raw[];
  double* s_key = reinterpret_cast<double*>(lx_raw);                    /* [LX_MAX_POINTS] each */
  double* s_a = s_key + LX_MAX_POINTS;
  double* s_r = s_a + LX_MAX_POINTS;
  int* s_idx = reinterpret_cast<int*>(s_r + LX_MAX_POINTS);
  int* s_flag = s_idx + LX_MAX_POINTS;
  __shared__ int s_red[32];
  __shared__ int s_misc[4];
  const int t = threadIdx.x;
  /* work on the smallest power-of-two window that holds the payload: a 361-beam scan costs what it did with a
   * 1024-point limit */
  const int npad = lx_npad(n_pairs), items = npad / LX_PREP_THREADS;
  /* polar points with a return; order-preserving compaction, one 1024-point chunk at a time */
  int np = 0;
  for (int c = 0; c < items; ++c) {
    const int i = c * LX_PREP_THREADS + t;
    float rr = 0.f, ang = 0.f;
    int keep = 0;
    if (i < n_pairs) { rr = b.data[2 * i]; ang = b.data[2 * i + 1]; keep = ((double)rr > 0.05) ? 1 : 0; }
    int cnt = 0;
    const int pos = np + block_scan_flags<LX_PREP_THREADS>(keep, s_red, &cnt);
    if (keep) { s_a[pos] = (double)ang - LX_MPI; s_r[pos] = (double)rr; }
    np += cnt;
  }
  __syncthreads();
  for (int i = t; i < npad; i += LX_PREP_THREADS) { s_key[i] = (i < np) ? s_a[i] + LX_PI : INFINITY; s_idx[i] = i; }
  __syncthreads();
  /* already sorted (the usual case for a rotating scanner)? */
  int inv = 0;
  for (int i = t; i + 1 < np; i += LX_PREP_THREADS) inv += (s_key[i + 1] < s_key[i]) ? 1 : 0;
  const int ninv = block_sum_int<LX_PREP_THREADS>(inv, s_red);
  if (ninv > 0) {
    /* bitonic sort of (key, original index): the index breaks ties, i.e. the sort is stable */
    for (int k = 2; k <= npad; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = t; i < npad; i += LX_PREP_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const bool up = (i & k) == 0;
            const double ka = s_key[i], kb = s_key[ixj];
            const int ia = s_idx[i], ib = s_idx[ixj];
            const bool a_after_b = (ka > kb) || (ka == kb && ia > ib);
            if (a_after_b == up) { s_key[i] = kb; s_key[ixj] = ka; s_idx[i] = ib; s_idx[ixj] = ia; }
          }
        }
        __syncthreads();
      }
    double na[LX_ITEMS], nr[LX_ITEMS];
#pragma unroll
    for (int c = 0; c < LX_ITEMS; ++c) {
      const int i = c * LX_PREP_THREADS + t;
      na[c] = 0.0; nr[c] = 0.0;
      if (c < items && i < np) { na[c] = s_a[s_idx[i]]; nr[c] = s_r[s_idx[i]]; }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < LX_ITEMS; ++c) {
      const int i = c * LX_PREP_THREADS + t;
      if (i < np) { s_a[i] = na[c]; s_r[i] = nr[c]; }
    }
    __syncthreads();
  }
  /* segmentation: split between neighbours more than 0.5 m apart */
  for (int pass = 0; pass < 2; ++pass) {
    if (t == 0) s_misc[0] = 0;
    __syncthreads();
    for (int i = t; i < npad; i += LX_PREP_THREADS) {
      int f = 0;
      if (i >= 1 && i < np) {
        const double r0 = s_r[i - 1], r1 = s_r[i];
        const double dist = sqrt(r0 * r0 + r1 * r1 - 2 * r0 * r1 * cos(s_a[i] - s_a[i - 1]));
        f = (dist > 0.5) ? 1 : 0;
      }
      s_flag[i] = f;
      if (f && pass == 0) atomicMax(&s_misc[0], i);                    /* split.back() */
    }
    __syncthreads();
    if (pass == 1) break;
    const int last = s_misc[0];
    if (last == 0) break;                                              /* no break in the scan: one segment, no rotation */
    double na[LX_ITEMS], nr[LX_ITEMS];
#pragma unroll
    for (int c = 0; c < LX_ITEMS; ++c) {
      const int i = c * LX_PREP_THREADS + t;
      na[c] = 0.0; nr[c] = 0.0;
      if (c < items && i < np) { const int src = (i + last) % np; na[c] = s_a[src]; nr[c] = s_r[src]; }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < LX_ITEMS; ++c) {
      const int i = c * LX_PREP_THREADS + t;
      if (i < np) { s_a[i] = na[c]; s_r[i] = nr[c]; }
    }
    __syncthreads();
  }
  /* segment table and point arrays */
  int nseg = 0;
  for (int c = 0; c < items; ++c) {
    const int i = c * LX_PREP_THREADS + t;
    const int is_start = (i < np) && (i == 0 || s_flag[i]);
    int cnt = 0;
    const int spos = nseg + block_scan_flags<LX_PREP_THREADS>(is_start, s_red, &cnt);
    if (is_start) b.seg[1 + spos] = i;
    nseg += cnt;
  }
  if (t == 0) { b.seg[0] = (np >= 2) ? nseg : 0; b.seg[1 + nseg] = np; *b.count = 0; }
  for (int i = t; i < npad; i += LX_PREP_THREADS) {
    if (i < np) {
      const double a = s_a[i], r = s_r[i];
      const double c = cos(a), sn = sin(a);
      b.a[i] = a; b.r[i] = r; b.ca[i] = c; b.sa[i] = sn;
      b.X[i] = c * r; b.Y[i] = sn * r;                                  /* polar2descart, lineFitting.cpp:157-168 */
    }
    b.leaf[i].valid = 0;
  }
}

/* ------------------------------------------------------------------------------------------------ */
struct LxFit { double alfa_raw, alfa, r, bb, m, SX, SY; };

/* line::line(alfa_deg, r) applied to the fit's result (lineFitting.cpp:17-23, 302) */
__device__ __forceinline__ void lx_make_line(LxFit& f, double alfa_raw, double r) {
  f.alfa_raw = alfa_raw;
  f.alfa = (alfa_raw * 180 / LX_PI) * (LX_PI / 180);
  f.r = r;
  f.bb = f.r / sin(f.alfa);
  f.m = -1 / tan(f.alfa);
}
__device__ __forceinline__ void lx_canonical(double& alfa, double& r) {      /* LineAlap, lineFitting.cpp:357-367 */
  if (r < 0) { r = fabs(r); alfa = (alfa < 0) ? LX_PI + alfa : -LX_PI + alfa; }
}
__device__ __forceinline__ double lx_alfanorm(double a) {
  if (a > LX_PI) return a - 2 * LX_PI;
  if (a < -LX_PI) return a + 2 * LX_PI;
  return a;
}
__device__ __forceinline__ double lx_len(double x, double y) {               /* Vec2::Lenght, vec2.cpp:26-34 */
  const double t = x * x + y * y;
  return (t > 0) ? sqrt(t) : 0;
}
/* FirstPoint / EndPoint, simplifyPath.cpp:59-105 */
__device__ void lx_end_point(const double* a, const double* r, int n, const LxFit& l, bool first_end, double& oa, double& orr) {
  const int e = first_end ? 0 : n - 1;
  if (n < 4) { oa = a[e]; orr = r[e]; return; }
  const double fr = l.r / (cos(a[n / 2] - l.alfa));
  const double er = l.r / (cos(a[e] - l.alfa));
  const double fa = a[n / 2], ea = a[e];
  const double Px = cos(a[e]) * r[e], Py = sin(a[e]) * r[e];
  const double fx = cos(fa) * fr, fy = sin(fa) * fr, ex = cos(ea) * er, ey = sin(ea) * er;
  const double FEx = ex - fx, FEy = ey - fy, FPx = Px - fx, FPy = Py - fy;
  const double lenFE = lx_len(FEx, FEy);
  const double nx = FEx / lenFE, ny = FEy / lenFE;
  const double skal = FEx * FPx + FEy * FPy;
  const double cs = skal / (lx_len(FEx, FEy) * lx_len(FPx, FPy));
  const double sc = cs * lx_len(FPx, FPy);
  const double Nx = fx + nx * sc, Ny = fy + ny * sc;
  orr = sqrt(Nx * Nx + Ny * Ny);
  oa = atan2(Ny, Nx);
}

/* One thread block per segment: simplifyPath::simplifyWithRDP (simplifyPath.cpp:108-177) with an explicit stack. */
__global__ void __launch_bounds__(LX_SEG_THREADS) k_lx_segments(LxBuffers b) {
  extern __shared__ unsigned char lx_raw[];
  double* s_a = reinterpret_cast<double*>(lx_raw);                      /* [LX_MAX_POINTS] each: a segment may hold every point */
  double* s_r = s_a + LX_MAX_POINTS;
  double* s_X = s_r + LX_MAX_POINTS;
  double* s_Y = s_X + LX_MAX_POINTS;
  int* s_stack = reinterpret_cast<int*>(s_Y + LX_MAX_POINTS);          /* [2 * LX_MAX_POINTS] */
  __shared__ double s_red[4 * (LX_SEG_THREADS / 32)];
  __shared__ double s_best[LX_SEG_THREADS / 32];
  __shared__ int s_besti[LX_SEG_THREADS / 32];
  __shared__ LxFit s_fit;
  __shared__ int s_dec[2];
  const int nseg = b.seg[0];
  const int t = threadIdx.x;
  for (int sg = blockIdx.x; sg < nseg; sg += gridDim.x) {
    const int g0 = b.seg[1 + sg], g1 = b.seg[2 + sg];
    const int P = g1 - g0;
    __syncthreads();
    for (int i = t; i < P; i += LX_SEG_THREADS) {
      s_a[i] = b.a[g0 + i]; s_r[i] = b.r[g0 + i]; s_X[i] = b.X[g0 + i]; s_Y[i] = b.Y[g0 + i];
    }
    int sp = 0;
    if (t == 0) { s_stack[0] = 0; s_stack[1] = P; }
    sp = 1;
    __syncthreads();
    while (sp > 0) {
      const int lo = s_stack[2 * (sp - 1)], hi = s_stack[2 * (sp - 1) + 1];
      sp -= 1;
      const int n = hi - lo;
      __syncthreads();
      if (n < 2) continue;                                              /* base case 1 */
      /* ---- fit, lineFitting.cpp:267-304 in closed form (see the header) ---- */
      double sx = 0.0, sy = 0.0;
      for (int i = lo + t; i < hi; i += LX_SEG_THREADS) { sx += s_X[i]; sy += s_Y[i]; }
      double r2[2] = {sx, sy};
      block_sum_n<LX_SEG_THREADS, 2>(r2, s_red);
      const double SX = r2[0], SY = r2[1];
      const double mx = SX / n, my = SY / n;
      double qxy = 0.0, qxx = 0.0, qyy = 0.0;
      for (int i = lo + t; i < hi; i += LX_SEG_THREADS) {
        const double dx = s_X[i] - mx, dy = s_Y[i] - my;
        qxy += dx * dy; qxx += dx * dx; qyy += dy * dy;
      }
      double r3[3] = {qxy, qxx, qyy};
      block_sum_n<LX_SEG_THREADS, 3>(r3, s_red);
      const double Sxy = r3[0], Sxx = r3[1], Syy = r3[2];
      if (t == 0) {
        const double alfa_raw = 0.5 * atan2(-2.0 * Sxy, -(Sxx - Syy));
        const double sum_r = cos(alfa_raw) * SX + sin(alfa_raw) * SY;
        lx_make_line(s_fit, alfa_raw, sum_r / n);
        s_fit.SX = SX; s_fit.SY = SY;
      }
      __syncthreads();
      const LxFit L = s_fit;
      /* ---- split test, simplifyPath.cpp:125-156 ---- */
      double q_di = 0.0, q_var = 0.0, q_t = 0.0;
      double best = -1.0; int besti = 0;
      const double fx = s_X[lo], fy = L.bb + s_X[lo] * L.m;              /* residual_error: chord of the FITTED line */
      const double lx = s_X[hi - 1], ly = L.bb + s_X[hi - 1] * L.m;
      const double dx = lx - fx, dy = ly - fy;
      const double dn = sqrt(dx * dx + dy * dy);
      const double cx0 = s_X[lo], cy0 = s_Y[lo];                        /* findMaximumDistance: chord first -> last point */
      const double ex = s_X[hi - 1] - cx0, ey = s_Y[hi - 1] - cy0;
      const double en = sqrt(ex * ex + ey * ey);
      for (int i = lo + t; i < hi; i += LX_SEG_THREADS) {
        const double cc = cos(s_a[i] - L.alfa);
        q_di += fabs(cc) * 2 * 0.01 / (sqrt(2 * LX_PI));
        q_var += cc * cc * 0.01 * 0.01 * ((LX_PI - 2) / LX_PI);
        if (i > lo) {
          const double px = s_X[i] - fx, py = s_Y[i] - fy;
          q_t += fabs(px * dy - dx * py) / dn;
          const double qx = s_X[i] - cx0, qy = s_Y[i] - cy0;
          const double dist = fabs(qx * ey - ex * qy) / en;
          if (dist > best) { best = dist; besti = i; }                 /* ascending i per thread: first maximum kept */
        }
      }
      double r4[3] = {q_di, q_var, q_t};
      block_sum_n<LX_SEG_THREADS, 3>(r4, s_red);
      const double sum_di = r4[0], sum_var = sqrt(r4[1]), tres = r4[2];
      /* arg max with the reference's tie rule (strict >, ascending index): larger distance, then smaller index */
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && ob >= 0.0 && oi < besti)) { best = ob; besti = oi; }
      }
      __syncthreads();
      if ((t & 31) == 0) { s_best[t >> 5] = best; s_besti[t >> 5] = besti; }
      __syncthreads();
      if (t == 0) {
        double bb = -1.0; int bi = 0;
        for (int w = 0; w < LX_SEG_THREADS / 32; ++w)
          if (s_best[w] > bb || (s_best[w] == bb && bb >= 0.0 && s_besti[w] < bi)) { bb = s_best[w]; bi = s_besti[w]; }
        const int index = (bb >= 0.0) ? bi - lo : 0;                   /* NaN distances leave index = 0 */
        int dec = 0;                                                   /* 0: leaf, 1: split, 2: drop (the reference would not terminate) */
        if (tres > sum_di + sum_var * 3) dec = (index <= 0 || index >= n) ? 2 : 1;
        s_dec[0] = dec; s_dec[1] = index;
        if (dec == 1) {                                                /* right part first: the left one is popped next */
          s_stack[2 * sp] = lo + index; s_stack[2 * sp + 1] = hi;
          s_stack[2 * sp + 2] = lo; s_stack[2 * sp + 3] = lo + index;
        }
      }
      __syncthreads();
      const int dec = s_dec[0];
      if (dec == 1) { sp += 2; continue; }
      if (dec == 2) continue;
      /* ---- leaf: covariance (lineFitting.cpp:379-450), end points, SetEndPoints ---- */
      double base_alfa = L.alfa, base_r = L.r;
      lx_canonical(base_alfa, base_r);
      if (base_alfa < 0) base_alfa = base_alfa + 2 * LX_PI;
      const double eps = 0.000001;
      const double cxv = 0.01 * 0.01 * 1.5;
      double q0 = 0.0, q3 = 0.0;
      for (int k = lo + t; k < hi; k += LX_SEG_THREADS) {
        const double rk = s_r[k] + eps;
        const double Xk = b.ca[g0 + k] * rk, Yk = b.sa[g0 + k] * rk;
        double ax = 0.0, ay = 0.0;
        for (int j = lo; j < hi; ++j) { ax += (j == k) ? Xk : s_X[j]; ay += (j == k) ? Yk : s_Y[j]; }
        const double ux = ax / n, uy = ay / n;
        double vxy = 0.0, vxx = 0.0, vyy = 0.0;
        for (int j = lo; j < hi; ++j) {
          const double ddx = ((j == k) ? Xk : s_X[j]) - ux, ddy = ((j == k) ? Yk : s_Y[j]) - uy;
          vxy += ddx * ddy; vxx += ddx * ddx; vyy += ddy * ddy;
        }
        const double ar = 0.5 * atan2(-2.0 * vxy, -(vxx - vyy));
        LxFit e;
        lx_make_line(e, ar, (cos(ar) * ax + sin(ar) * ay) / n);
        double ea = e.alfa, er = e.r;
        lx_canonical(ea, er);
        const double al = (ea < 0) ? ea + 2 * LX_PI : ea;
        const double F0 = lx_alfanorm(al - base_alfa) / eps;
        const double F1 = (er - base_r) / eps;
        q0 += (F0 * cxv) * F0;                                         /* (F C_x) F^T, diagonal C_x (:443-444) */
        q3 += (F1 * cxv) * F1;
      }
      double r5[2] = {q0, q3};
      block_sum_n<LX_SEG_THREADS, 2>(r5, s_red);
      const double C0 = r5[0], C3 = r5[1];
      if (t == 0) {
        LxLeaf lf;
        lf.alfa = L.alfa; lf.r = L.r; lf.c0 = C0; lf.c3 = C3;
        double ia0, ir0, ia1, ir1;
        lx_end_point(s_a + lo, s_r + lo, n, L, true, ia0, ir0);
        lx_end_point(s_a + lo, s_r + lo, n, L, false, ia1, ir1);
        ir0 = L.r / (cos(ia0 - L.alfa)); ir1 = L.r / (cos(ia1 - L.alfa));   /* line::SetEndPoints, lineFitting.cpp:44-51 */
        ia0 = ia0 + LX_PI; ia1 = ia1 + LX_PI;
        ia0 = ia0 > LX_PI ? ia0 - 2.0 * LX_PI : ia0;
        ia1 = ia1 > LX_PI ? ia1 - 2.0 * LX_PI : ia1;
        lf.ia0 = ia0; lf.ir0 = ir0; lf.ia1 = ia1; lf.ir1 = ir1;
        lf.valid = 1; lf.pad = 0;
        b.leaf[g0 + lo] = lf;
      }
    }
  }
}

/* LineConversion (lineFitting.cpp:586-638) + main.cpp:66-69, compacted in leaf (= reference) order. */
__global__ void __launch_bounds__(LX_PREP_THREADS) k_lx_finish(LxBuffers b, int max_lines, int n_pairs) {
  __shared__ int s_red[32];
  const int t = threadIdx.x;
  const int items = lx_npad(n_pairs) / LX_PREP_THREADS;
  int total = 0;
  for (int c = 0; c < items; ++c) {
    const LxLeaf lf = b.leaf[c * LX_PREP_THREADS + t];
    int keep = 0;
    if (lf.valid) {
      keep = 1;
      if (lf.c0 < 0 || lf.c3 < 0) keep = 0;
      else if (isnan(lf.c0) || isnan(lf.c3)) keep = 0;
      else if (lf.alfa == 0 && lf.r == 0) keep = 0;
      else if (lf.c0 > 0.01) keep = 0;
    }
    int cnt = 0;
    const int pos = total + block_scan_flags<LX_PREP_THREADS>(keep, s_red, &cnt);
    total += cnt;
    if (keep && pos < max_lines) {
      double alfa = lf.alfa, r = lf.r;
      lx_canonical(alfa, r);
      alfa += LX_MPI;                                                  /* main.cpp:67-68 */
      alfa = alfa > LX_MPI ? alfa - 2.0 * LX_MPI : alfa;
      double* o = b.out + 10 * pos;
      o[0] = alfa; o[1] = r; o[2] = lf.c0; o[3] = 0.0; o[4] = 0.0; o[5] = lf.c3;
      o[6] = lf.ia0; o[7] = lf.ir0; o[8] = lf.ia1; o[9] = lf.ir1;
      b.z[2 * pos] = alfa; b.z[2 * pos + 1] = r;
      b.R[4 * pos] = lf.c0; b.R[4 * pos + 1] = 0.0; b.R[4 * pos + 2] = 0.0; b.R[4 * pos + 3] = lf.c3;
    }
  }
  if (t == 0) *b.count = total;
}

/* lineprovider/main.cpp:60-84 (Transform): the two end points of every line of the last extraction from the robot frame
 * to the world frame, ex = (cos theta, sin theta), ey = (cos, sin)(theta + PI/2) with the reference's truncated PI
 * (lineFitting.h:12) and its wrap at M_PI, p = ex p.x + ey p.y + pose, every product and sum rounded on its own as the
 * reference's Vec2 operators do; then astar/main.cpp:44-73 (lines_cb): the planner takes the FLOATS of the `lines_1`
 * message times 100 (centimetres) -- a float product.  One thread per line; out: 4 floats per line. */
__global__ void k_lx_world(LxBuffers b, int max_lines, double px, double py, double theta, float scale, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(*b.count, max_lines);
  if (i >= n) return;
  const double exx = cos(theta), exy = sin(theta);
  double t = theta + 3.14159265 / 2;
  t = t > LX_MPI ? t - 2.0 * LX_MPI : t;
  const double eyx = cos(t), eyy = sin(t);
  const double* o = b.out + 10 * i;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double a = o[6 + 2 * k], r = o[7 + 2 * k];
    const double fx = __dmul_rn(cos(a), r), fy = __dmul_rn(sin(a), r);              /* polar2descart, lineFitting.cpp:170-176 */
    const double wx = __dadd_rn(__dadd_rn(__dmul_rn(exx, fx), __dmul_rn(eyx, fy)), px);
    const double wy = __dadd_rn(__dadd_rn(__dmul_rn(exy, fx), __dmul_rn(eyy, fy)), py);
    out[4 * i + 2 * k] = __fmul_rn((float)wx, scale);
    out[4 * i + 2 * k + 1] = __fmul_rn((float)wy, scale);
  }
}

}  // namespace

/* ------------------------------------------------------------------------------------------------ */
static const size_t kPrepSmem = (size_t)LX_MAX_POINTS * (3 * sizeof(double) + 2 * sizeof(int));   /* 128 KB */
static const size_t kSegSmem = (size_t)LX_MAX_POINTS * (4 * sizeof(double) + 2 * sizeof(int));    /* 160 KB */

struct ekf_lx {
  int device, max_lines;
  cudaStream_t stream;
  LxBuffers b;
  float* d_data; float* h_data;
  double* h_out; int* h_count;
  float* d_seg; float* h_seg;         /* world-frame end points of the last extraction's lines (ekf_lx_world_segments) */
  long long launches;
  char err[256];
};

#define LXCU(call)                                                                               \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      snprintf(lx->err, sizeof lx->err, "%s:%d %.120s: %s", "ekf_lines.cu", __LINE__, #call, cudaGetErrorString(e_)); \
      return EKF_ECUDA;                                                                          \
    }                                                                                            \
  } while (0)

extern "C" {

int ekf_lx_create(ekf_lx** out, int device, int max_lines) {
  if (!out || max_lines < 1) return EKF_EINVAL;
  ekf_lx* lx = new ekf_lx();
  memset(lx, 0, sizeof *lx);
  lx->device = device; lx->max_lines = max_lines;
  *out = lx;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    snprintf(lx->err, sizeof lx->err, "no CUDA device: libekfcuda has no CPU fallback");
    return EKF_ECUDA;
  }
  LXCU(cudaSetDevice(device));
  LXCU(cudaStreamCreateWithFlags(&lx->stream, cudaStreamNonBlocking));
  const size_t np = LX_MAX_POINTS;
  LXCU(cudaMalloc(&lx->d_data, 2 * np * sizeof(float)));
  LXCU(cudaMallocHost(&lx->h_data, 2 * np * sizeof(float)));
  double* pts = 0;
  LXCU(cudaMalloc(&pts, 6 * np * sizeof(double)));
  lx->b.a = pts; lx->b.r = pts + np; lx->b.X = pts + 2 * np; lx->b.Y = pts + 3 * np; lx->b.ca = pts + 4 * np; lx->b.sa = pts + 5 * np;
  LXCU(cudaMalloc(&lx->b.seg, (np + 2) * sizeof(int)));
  LXCU(cudaMalloc(&lx->b.leaf, np * sizeof(LxLeaf)));
  LXCU(cudaMalloc(&lx->b.out, 10 * (size_t)max_lines * sizeof(double)));
  LXCU(cudaMalloc(&lx->b.z, 2 * (size_t)max_lines * sizeof(double)));
  LXCU(cudaMalloc(&lx->b.R, 4 * (size_t)max_lines * sizeof(double)));
  LXCU(cudaMalloc(&lx->b.count, sizeof(int)));
  LXCU(cudaMallocHost(&lx->h_out, 10 * (size_t)max_lines * sizeof(double)));
  LXCU(cudaMallocHost(&lx->h_count, sizeof(int)));
  LXCU(cudaMalloc(&lx->d_seg, 4 * (size_t)max_lines * sizeof(float)));
  LXCU(cudaMallocHost(&lx->h_seg, 4 * (size_t)max_lines * sizeof(float)));
  lx->b.data = lx->d_data;
  LXCU(cudaFuncSetAttribute(k_lx_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepSmem));
  LXCU(cudaFuncSetAttribute(k_lx_segments, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSegSmem));
  LXCU(cudaStreamSynchronize(lx->stream));
  return EKF_OK;
}

int ekf_lx_destroy(ekf_lx* lx) {
  if (!lx) return EKF_EINVAL;
  cudaSetDevice(lx->device);
  if (lx->stream) cudaStreamSynchronize(lx->stream);
  cudaFree(lx->d_data); cudaFreeHost(lx->h_data); cudaFree(lx->b.a); cudaFree(lx->b.seg); cudaFree(lx->b.leaf);
  cudaFree(lx->b.out); cudaFree(lx->b.z); cudaFree(lx->b.R); cudaFree(lx->b.count);
  cudaFreeHost(lx->h_out); cudaFreeHost(lx->h_count); cudaFree(lx->d_seg); cudaFreeHost(lx->h_seg);
  if (lx->stream) cudaStreamDestroy(lx->stream);
  delete lx;
  return EKF_OK;
}

const char* ekf_lx_last_error(const ekf_lx* lx) { return lx ? lx->err : "null ekf_lx"; }

static int lx_enqueue(ekf_lx* lx, const float* d_data, int n_pairs) {
  LxBuffers b = lx->b;
  b.data = d_data;
  k_lx_prepare<<<1, LX_PREP_THREADS, kPrepSmem, lx->stream>>>(b, n_pairs);
  LXCU(cudaGetLastError());
  k_lx_segments<<<148, LX_SEG_THREADS, kSegSmem, lx->stream>>>(b);
  LXCU(cudaGetLastError());
  k_lx_finish<<<1, LX_PREP_THREADS, 0, lx->stream>>>(b, lx->max_lines, n_pairs);
  LXCU(cudaGetLastError());
  lx->launches += 3;
  return EKF_OK;
}

int ekf_lx_extract(ekf_lx* lx, int n_pairs, const float* data, int* n_lines, double* lines) {
  if (!lx || n_pairs < 0 || n_pairs > LX_MAX_POINTS || (n_pairs > 0 && !data) || !n_lines) return EKF_EINVAL;
  LXCU(cudaSetDevice(lx->device));
  if (n_pairs > 0) memcpy(lx->h_data, data, 2 * (size_t)n_pairs * sizeof(float));
  if (n_pairs > 0) LXCU(cudaMemcpyAsync(lx->d_data, lx->h_data, 2 * (size_t)n_pairs * sizeof(float), cudaMemcpyHostToDevice, lx->stream));
  int rc = lx_enqueue(lx, lx->d_data, n_pairs);
  if (rc) return rc;
  LXCU(cudaMemcpyAsync(lx->h_count, lx->b.count, sizeof(int), cudaMemcpyDeviceToHost, lx->stream));
  LXCU(cudaMemcpyAsync(lx->h_out, lx->b.out, 10 * (size_t)lx->max_lines * sizeof(double), cudaMemcpyDeviceToHost, lx->stream));
  LXCU(cudaStreamSynchronize(lx->stream));
  *n_lines = *lx->h_count;
  const int nw = (*n_lines < lx->max_lines) ? *n_lines : lx->max_lines;
  if (lines && nw > 0) memcpy(lines, lx->h_out, 10 * (size_t)nw * sizeof(double));
  return EKF_OK;
}

int ekf_lx_extract_device(ekf_lx* lx, int n_pairs, const float* d_data, const double** d_z, const double** d_R, const int** d_count) {
  if (!lx || n_pairs < 0 || n_pairs > LX_MAX_POINTS || (n_pairs > 0 && !d_data)) return EKF_EINVAL;
  LXCU(cudaSetDevice(lx->device));
  int rc = lx_enqueue(lx, d_data, n_pairs);
  if (rc) return rc;
  if (d_z) *d_z = lx->b.z;
  if (d_R) *d_R = lx->b.R;
  if (d_count) *d_count = lx->b.count;
  return EKF_OK;
}

int ekf_lx_world_segments(ekf_lx* lx, const double pose[3], double scale, float* out, int max_out_lines, int* n_lines) {
  if (!lx || !pose || !out || max_out_lines < 0 || !n_lines) return EKF_EINVAL;
  LXCU(cudaSetDevice(lx->device));
  k_lx_world<<<(lx->max_lines + 127) / 128, 128, 0, lx->stream>>>(lx->b, lx->max_lines, pose[0], pose[1], pose[2], (float)scale, lx->d_seg);
  LXCU(cudaGetLastError());
  lx->launches += 1;
  LXCU(cudaMemcpyAsync(lx->h_count, lx->b.count, sizeof(int), cudaMemcpyDeviceToHost, lx->stream));
  LXCU(cudaMemcpyAsync(lx->h_seg, lx->d_seg, 4 * (size_t)lx->max_lines * sizeof(float), cudaMemcpyDeviceToHost, lx->stream));
  LXCU(cudaStreamSynchronize(lx->stream));
  int n = *lx->h_count < lx->max_lines ? *lx->h_count : lx->max_lines;
  *n_lines = n;
  if (n > max_out_lines) n = max_out_lines;
  if (n > 0) memcpy(out, lx->h_seg, 4 * (size_t)n * sizeof(float));
  return EKF_OK;
}

int ekf_lx_sync(ekf_lx* lx) {
  if (!lx) return EKF_EINVAL;
  LXCU(cudaSetDevice(lx->device));
  LXCU(cudaStreamSynchronize(lx->stream));
  return EKF_OK;
}

}  /* extern "C" */
