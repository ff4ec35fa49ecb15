/* ekf_batch.cu -- independent filter instances (Monte-Carlo batches, BASELINE.json configs[3]).
 *
 * One 128-thread block per filter; the whole Robot::localize (slam_ros/Robot.cpp:126-943) of that filter
 * runs inside one kernel launch.  Same design as the single big filter, scaled down to one CTA: the hot
 * state (rows 0..2 of P, the 2x2 diagonal blocks, y) and the scan's pending gains live in shared memory
 * (~23 KB, so four filters share an SM and hide each other's fp64 latency chains); the cold part of P stays
 * in HBM, is read where a gain needs a column (with the pending terms applied on the fly) and is brought up
 * to date by ONE read-modify-write pass over its upper triangle at the end of the scan: 8 n (n+1) bytes per
 * filter-scan however many lines match.  Filters never communicate; a multi-GPU batch is N independent
 * ekf_batch objects, one per device (slam_ros_b200/parallel.py deals filters round-robin).
 */
#include "../../include/ekf.h"
#include "ekf_internal.h"
#include "ekf_device.cuh"

#include <stdio.h>
#include <string.h>

#define EKFB_THREADS 128

struct EkfBatchState {
  double pose[3];
  int L;
  int sticky;
  int resets;
  int pad;
};

struct EkfBatchGeom {
  int B, cap, n, headroom;
  double gate, enc_noise;
};

namespace {

#define EKFB_SLOTS 8        /* pending rank-2 terms kept in shared memory before the cold part is swept */

/* Shared-memory footprint: hot state + pending gains only (the cold part of P never leaves HBM/L2 except for
 * the one read-modify-write sweep per scan), so several filters share an SM. */
struct BatchSmem {
  double* top;              /* [3][n]  rows 0..2 of P, upper authoritative */
  double* diag;             /* [cap][3] P[a,a], P[a,b], P[b,b] */
  double* y;                /* [n] */
  double2* K;               /* [EKFB_SLOTS][n] pending gains */
  double* S;                /* [EKFB_SLOTS][4] their innovation covariances (K S is recomputed: Robot.cpp:560) */
  int* ext;                 /* [m] */
  unsigned char* matched;   /* [cap] */
};
__host__ __device__ inline size_t batch_smem_bytes(int n, int cap, int m) {
  size_t d = (size_t)(3 * n + 3 * cap + n);
  d += d & 1;                                      /* double2 alignment of the pending gains */
  size_t b = d * sizeof(double);
  b += (size_t)EKFB_SLOTS * n * sizeof(double2) + (size_t)EKFB_SLOTS * 4 * sizeof(double);
  b += (size_t)m * sizeof(int) + (size_t)cap;
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ bool b_is_hot(int r, int q) { return r <= 2 || q == r || (q == r + 1 && (r & 1)); }
__device__ __forceinline__ double b_hot(const BatchSmem& sm, int n, int r, int q) {
  if (r <= 2) return sm.top[r * n + q];
  const int j = (r - 3) >> 1;
  return (r & 1) ? sm.diag[3 * j + (q - r)] : sm.diag[3 * j + 2];
}
/* (K S)[r] of pending term i, Robot.cpp:560 (NN, zero skip) */
__device__ __forceinline__ double2 b_ks(const BatchSmem& sm, int n, int i, int r) {
  const double2 k = sm.K[i * n + r];
  const double* S = sm.S + 4 * i;
  double s0 = 0.0, s1 = 0.0;
  axpy_skip(s0, k.x, S[0]); axpy_skip(s1, k.x, S[1]);
  axpy_skip(s0, k.y, S[2]); axpy_skip(s1, k.y, S[3]);
  return make_double2(s0, s1);
}
/* current value of a cold upper element: HBM value minus the pending terms, in order */
__device__ __forceinline__ double b_cold(const BatchSmem& sm, const double* __restrict__ Pf, int n, int r, int q, int np) {
  double p = Pf[(size_t)r * n + q];
  for (int i = 0; i < np; ++i) p = sub_rank2(p, b_ks(sm, n, i, r), sm.K[i * n + q]);
  return p;
}
/* fold the pending terms into the cold upper triangle of the live part (one read-modify-write pass) */
__device__ void b_sweep(const BatchSmem& sm, double* __restrict__ Pf, int n, int nl, int np) {
  if (np <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = 3 + warp; r < nl; r += nw) {
    double2 ks[EKFB_SLOTS];
#pragma unroll
    for (int i = 0; i < EKFB_SLOTS; ++i) if (i < np) ks[i] = b_ks(sm, n, i, r);
    double* Pr = Pf + (size_t)r * n;
    const int q0 = (r & 1) ? r + 2 : r + 1;            /* skip the 2x2 diagonal block (hot) */
    for (int q = q0 + lane; q < nl; q += 32) {
      double p = Pr[q];
#pragma unroll
      for (int i = 0; i < EKFB_SLOTS; ++i) if (i < np) p = sub_rank2(p, ks[i], sm.K[i * n + q]);
      Pr[q] = p;
    }
  }
}

/* Robot::localize for filter blockIdx.x */
__global__ void __launch_bounds__(EKFB_THREADS, 8) k_batch_scan(EkfBatchGeom g, double* __restrict__ Yg,
                                                                double* __restrict__ Pg, EkfBatchState* __restrict__ Sg,
                                                                const double* __restrict__ U, const double* __restrict__ Z,
                                                                const double* __restrict__ Rm, int m, int* __restrict__ Jout) {
  extern __shared__ double smem[];
  const int n = g.n, tid = threadIdx.x, nt = blockDim.x;
  const int f = blockIdx.x;
  BatchSmem sm;
  sm.top = smem;
  sm.diag = sm.top + 3 * n;
  sm.y = sm.diag + 3 * g.cap;
  sm.K = reinterpret_cast<double2*>(sm.y + n + ((3 * n + 3 * g.cap + n) & 1));
  sm.S = reinterpret_cast<double*>(sm.K + (size_t)EKFB_SLOTS * n);
  sm.ext = reinterpret_cast<int*>(sm.S + EKFB_SLOTS * 4);
  sm.matched = reinterpret_cast<unsigned char*>(sm.ext + m);
  __shared__ int s_min[EKFB_THREADS / 32];
  __shared__ int s_best, s_ne, s_nmatch, s_L, s_stop;
  __shared__ double s_xpre[3], s_pose[3], s_cs[2];
  __shared__ Gate sG;

  double* Pf = Pg + (size_t)f * n * n;
  double* yf = Yg + (size_t)f * n;
  EkfBatchState* st = Sg + f;
  const double* u = U + 3 * (size_t)f;
  const double* z = Z + 2 * (size_t)m * f;
  const double* R = Rm + 4 * (size_t)m * f;
  int* jout = Jout ? Jout + (size_t)m * f : 0;

  if (tid == 0) {
    st->sticky = 0;                           /* status reports this scan only */
    s_L = st->L; s_ne = 0; s_nmatch = 0;
    s_pose[0] = st->pose[0]; s_pose[1] = st->pose[1]; s_pose[2] = st->pose[2];
  }
  __syncthreads();
  int L = s_L;
  int nl = 3 + 2 * L;
  for (int i = tid; i < 3 * n; i += nt) { const int r = i / n, q = i % n; sm.top[i] = (q < nl) ? Pf[(size_t)r * n + q] : 0.0; }
  for (int j = tid; j < g.cap; j += nt) {
    const int a = 3 + 2 * j;
    if (j < L) { sm.diag[3 * j] = Pf[(size_t)a * n + a]; sm.diag[3 * j + 1] = Pf[(size_t)a * n + a + 1]; sm.diag[3 * j + 2] = Pf[(size_t)(a + 1) * n + a + 1]; }
    sm.matched[j] = 0;
  }
  for (int i = tid; i < n; i += nt) sm.y[i] = yf[i];
  __syncthreads();

  /* ---- prediction, Robot.cpp:130-258 ---- */
  const double u0 = u[0], u2 = u[2];
  const double ang = add_rn(s_pose[2], __ddiv_rn(u2, 2.0));
  const double ca = cos(ang), sa = sin(ang);
  const double F02 = mul_rn(-u0, sa), F12 = mul_rn(u0, ca);
  for (int q = 3 + tid; q < nl; q += nt) {                            /* :242 rows 0,1 */
    const double p2 = sm.top[2 * n + q];
    double a0 = add_rn(0.0, sm.top[q]); axpy_skip(a0, F02, p2);
    double a1 = add_rn(0.0, sm.top[n + q]); axpy_skip(a1, F12, p2);
    sm.top[q] = a0; sm.top[n + q] = a1;
  }
  if (tid == 0) {
    double A[3][3], T[3][3], Pn[3][3];
    for (int i = 0; i < 3; ++i) for (int k = i; k < 3; ++k) { A[i][k] = sm.top[i * n + k]; A[k][i] = A[i][k]; }
    for (int j = 0; j < 3; ++j) {
      double r0 = 0.0, r1 = 0.0, r2 = 0.0;
      axpy_skip(r0, 1.0, A[0][j]); axpy_skip(r1, 1.0, A[1][j]);
      axpy_skip(r0, F02, A[2][j]); axpy_skip(r1, F12, A[2][j]); axpy_skip(r2, 1.0, A[2][j]);
      T[0][j] = r0; T[1][j] = r1; T[2][j] = r2;
    }
    for (int i = 0; i < 3; ++i) {
      double c0 = 0.0; c0 = add_rn(c0, mul_rn(T[i][0], 1.0)); c0 = add_rn(c0, mul_rn(T[i][2], F02));
      double c1 = 0.0; c1 = add_rn(c1, mul_rn(T[i][1], 1.0)); c1 = add_rn(c1, mul_rn(T[i][2], F12));
      Pn[i][0] = add_rn(0.0, c0); Pn[i][1] = add_rn(0.0, c1); Pn[i][2] = T[i][2];
    }
    const double Fu[3][3] = {{ca, 0.0, __ddiv_rn(mul_rn(-u0, sa), 2.0)},
                             {sa, 1.0, __ddiv_rn(mul_rn(u0, ca), 2.0)},
                             {0.0, 0.0, 1.0}};
    const double qf = add_rn(__ddiv_rn(-1.0, add_rn(1.0, fabs(u0))), 1.0);
    const double Q[3] = {mul_rn(g.enc_noise, qf), mul_rn(mul_rn(2.0, g.enc_noise), qf), mul_rn(g.enc_noise, qf)};
    double FQ[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < 3; ++k)
      for (int i = 0; i < 3; ++i) {
        const double t = mul_rn(1.0, Fu[i][k]);
        if (t != 0.0)
          for (int j = 0; j < 3; ++j) FQ[i][j] = add_rn(FQ[i][j], mul_rn(t, (j == k) ? Q[k] : 0.0));
      }
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) {
        double t = 0.0;
        for (int k = 0; k < 3; ++k) t = add_rn(t, mul_rn(FQ[i][k], Fu[j][k]));
        sm.top[i * n + j] = add_rn(Pn[i][j], add_rn(0.0, mul_rn(1.0, t)));
      }
    s_xpre[0] = add_rn(s_pose[0], mul_rn(u0, ca));
    s_xpre[1] = add_rn(s_pose[1], mul_rn(u0, sa));
    s_xpre[2] = add_rn(s_pose[2], u2);
  }
  __syncthreads();

  /* ---- the observed lines in order, Robot.cpp:298-645 ---- */
  int np = 0;                                  /* pending terms (uniform across the block) */
  for (int i = 0; i < m; ++i) {
    const double z0 = z[2 * i], z1 = z[2 * i + 1];
    const double Rl[4] = {R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]};
    const double xp[3] = {s_xpre[0], s_xpre[1], s_xpre[2]};
    int cand = EKF_NO_MATCH;
    Gate G;
    for (int j = tid; j < L; j += nt) {
      if (sm.matched[j] || cand != EKF_NO_MATCH) continue;
      const int a = 3 + 2 * j, bb = a + 1;
      double Cm[5][5];
      for (int r = 0; r < 3; ++r) {
        for (int q = r; q < 3; ++q) { Cm[r][q] = sm.top[r * n + q]; Cm[q][r] = Cm[r][q]; }
        Cm[r][3] = Cm[3][r] = sm.top[r * n + a]; Cm[r][4] = Cm[4][r] = sm.top[r * n + bb];
      }
      Cm[3][3] = sm.diag[3 * j]; Cm[3][4] = Cm[4][3] = sm.diag[3 * j + 1]; Cm[4][4] = sm.diag[3 * j + 2];
      Gate Gj;
      gate_from_block(Cm, sm.y[a], sm.y[bb], xp, z0, z1, Rl, Gj);
      if (Gj.singular) atomicOr(&st->sticky, EKF_STICKY_SINGULAR);
      else if (!(sqrt(fabs(Gj.d2)) > g.gate)) { cand = j; G = Gj; }
    }
    const int mine = cand;
    cand = __reduce_min_sync(0xffffffffu, cand);
    if ((tid & 31) == 0) s_min[tid >> 5] = cand;
    __syncthreads();
    if (tid < 32) {
      int v = (tid < EKFB_THREADS / 32) ? s_min[tid] : EKF_NO_MATCH;
      v = __reduce_min_sync(0xffffffffu, v);
      if (tid == 0) {
        s_best = v;
        if (v == EKF_NO_MATCH) { sm.ext[s_ne++] = i; if (jout) jout[i] = -1; }
      }
    }
    __syncthreads();
    const int jb = s_best;
    if (jb == EKF_NO_MATCH) continue;
    if (mine == jb) sG = G;                     /* the winner publishes its gate record (evaluated once) */
    if (np == EKFB_SLOTS) {                     /* pending list full: fold it into the cold part first */
      __syncthreads();
      b_sweep(sm, Pf, n, nl, np);
      np = 0;
    }
    __syncthreads();
    const int a = 3 + 2 * jb, bb = a + 1;
    for (int r = tid; r < nl; r += nt) {                              /* :516-560 */
      double p0, p1, p2;
      if (r <= 2) { p0 = sm.top[min(r, 0) * n + max(r, 0)]; p1 = sm.top[min(r, 1) * n + max(r, 1)]; p2 = sm.top[min(r, 2) * n + max(r, 2)]; }
      else { p0 = sm.top[r]; p1 = sm.top[n + r]; p2 = sm.top[2 * n + r]; }
      const int lo_a = min(r, a), hi_a = max(r, a), lo_b = min(r, bb), hi_b = max(r, bb);
      const double pa = b_is_hot(lo_a, hi_a) ? b_hot(sm, n, lo_a, hi_a) : b_cold(sm, Pf, n, lo_a, hi_a, np);
      const double pb = b_is_hot(lo_b, hi_b) ? b_hot(sm, n, lo_b, hi_b) : b_cold(sm, Pf, n, lo_b, hi_b, np);
      double2 Kr, KSr;
      gain_row(sG, p0, p1, p2, pa, pb, Kr, KSr);
      sm.K[np * n + r] = Kr;
    }
    if (tid < 4) sm.S[4 * np + tid] = sG.S[tid];
    __syncthreads();
    {                                                                 /* :564-602 on the hot elements */
      const double2 ks0 = b_ks(sm, n, np, 0), ks1 = b_ks(sm, n, np, 1), ks2 = b_ks(sm, n, np, 2);
      const double v0 = sG.v[0], v1 = sG.v[1];
      for (int q = 3 + tid; q < nl; q += nt) {
        const double2 kq = sm.K[np * n + q];
        sm.top[q] = sub_rank2(sm.top[q], ks0, kq);
        sm.top[n + q] = sub_rank2(sm.top[n + q], ks1, kq);
        sm.top[2 * n + q] = sub_rank2(sm.top[2 * n + q], ks2, kq);
        const double2 ksq = b_ks(sm, n, np, q);
        const int jj = (q - 3) >> 1;
        if (q & 1) {
          sm.diag[3 * jj] = sub_rank2(sm.diag[3 * jj], ksq, kq);
          sm.diag[3 * jj + 1] = sub_rank2(sm.diag[3 * jj + 1], ksq, sm.K[np * n + q + 1]);
        } else {
          sm.diag[3 * jj + 2] = sub_rank2(sm.diag[3 * jj + 2], ksq, kq);
        }
        double t = 0.0;
        axpy_skip(t, kq.x, v0); axpy_skip(t, kq.y, v1);
        sm.y[q] = add_rn(sm.y[q], t);
      }
      if (tid == 0) {
        const double2 kk[3] = {sm.K[np * n], sm.K[np * n + 1], sm.K[np * n + 2]};
        const double2 ks[3] = {ks0, ks1, ks2};
        for (int r = 0; r < 3; ++r)
          for (int q = r; q < 3; ++q) sm.top[r * n + q] = sub_rank2(sm.top[r * n + q], ks[r], kk[q]);
        double yn[3];
        for (int r = 0; r < 3; ++r) {
          double t = 0.0;
          axpy_skip(t, kk[r].x, v0); axpy_skip(t, kk[r].y, v1);
          yn[r] = add_rn(s_xpre[r], t);
        }
        normalize_radian(yn[2]);
        for (int r = 0; r < 3; ++r) { sm.y[r] = yn[r]; s_pose[r] = yn[r]; s_xpre[r] = yn[r]; }
        sm.matched[jb] = 1; s_nmatch++;
        if (jout) jout[i] = jb;
      }
    }
    np += 1;
    __syncthreads();
  }

  /* ---- the one deferred sweep of the scan: cold upper triangle -= pending terms ---- */
  b_sweep(sm, Pf, n, nl, np);

  /* ---- Robot.cpp:702-716 ---- */
  if (tid == 0) {
    if (m == 0 || s_nmatch == 0) {
      sm.y[0] = s_xpre[0]; sm.y[1] = s_xpre[1]; sm.y[2] = s_xpre[2];
      double th = s_xpre[2];
      normalize_radian(th);
      s_pose[0] = s_xpre[0]; s_pose[1] = s_xpre[1]; s_pose[2] = th;
    }
    s_stop = 0;
  }
  __syncthreads();

  /* ---- augmentation in queue order, Robot.cpp:776-866 ---- */
  const int ne = s_ne;
  for (int e = 0; e < ne; ++e) {
    const int l = 3 + 2 * L;
    if (tid == 0) {
      if (L >= g.cap) { atomicOr(&st->sticky, EKF_STICKY_CAPACITY); s_stop = 1; }
      else {
        const int i = sm.ext[e];
        double alfa = z[2 * i], r = z[2 * i + 1];
        const double Rl[4] = {R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]};
        r = add_rn(r, add_rn(mul_rn(s_pose[0], cos(alfa)), mul_rn(s_pose[1], sin(alfa))));
        alfa = add_rn(alfa, s_pose[2]);
        const double cw = cos(alfa), sw = sin(alfa);
        const double Gx[2][3] = {{0.0, 0.0, 1.0}, {cw, sw, 0.0}};
        const double Gl[2][2] = {{1.0, 0.0}, {sub_rn(mul_rn(sm.y[1], cw), mul_rn(sm.y[0], sw)), 1.0}};
        normalize_radian(alfa);
        sm.y[l] = alfa; sm.y[l + 1] = r;
        s_cs[0] = cw; s_cs[1] = sw;
        double A[3][3];
        for (int ii = 0; ii < 3; ++ii) for (int kk = ii; kk < 3; ++kk) { A[ii][kk] = sm.top[ii * n + kk]; A[kk][ii] = A[ii][kk]; }
        double GP[2][3] = {{0, 0, 0}, {0, 0, 0}};
        for (int k = 0; k < 3; ++k)
          for (int ii = 0; ii < 2; ++ii) {
            const double t = mul_rn(1.0, Gx[ii][k]);
            if (t != 0.0) for (int jj = 0; jj < 3; ++jj) GP[ii][jj] = add_rn(GP[ii][jj], mul_rn(t, A[k][jj]));
          }
        double Pll[2][2];
        for (int ii = 0; ii < 2; ++ii)
          for (int jj = 0; jj < 2; ++jj) {
            double t = 0.0;
            for (int k = 0; k < 3; ++k) t = add_rn(t, mul_rn(GP[ii][k], Gx[jj][k]));
            Pll[ii][jj] = add_rn(0.0, mul_rn(1.0, t));
          }
        double GR[2][2] = {{0, 0}, {0, 0}};
        for (int k = 0; k < 2; ++k)
          for (int ii = 0; ii < 2; ++ii) {
            const double t = mul_rn(1.0, Gl[ii][k]);
            if (t != 0.0) for (int jj = 0; jj < 2; ++jj) GR[ii][jj] = add_rn(GR[ii][jj], mul_rn(t, Rl[k * 2 + jj]));
          }
        for (int ii = 0; ii < 2; ++ii)
          for (int jj = 0; jj < 2; ++jj) {
            double t = 0.0;
            for (int k = 0; k < 2; ++k) t = add_rn(t, mul_rn(GR[ii][k], Gl[jj][k]));
            Pll[ii][jj] = add_rn(Pll[ii][jj], add_rn(0.0, mul_rn(1.0, t)));
          }
        sm.diag[3 * L] = Pll[0][0]; sm.diag[3 * L + 1] = Pll[0][1]; sm.diag[3 * L + 2] = Pll[1][1];
      }
    }
    __syncthreads();
    if (s_stop) break;
    const double cw = s_cs[0], sw = s_cs[1];
    for (int k = tid; k < l; k += nt) {                               /* :856-860: the new landmark's column block */
      const double a0 = (k <= 0) ? sm.top[k * n + 0] : sm.top[0 * n + k];   /* P[0,k] through the upper storage */
      const double a1 = (k <= 1) ? sm.top[k * n + 1] : sm.top[1 * n + k];
      const double a2 = (k <= 2) ? sm.top[k * n + 2] : sm.top[2 * n + k];
      double r0 = 0.0, r1 = 0.0;
      axpy_skip(r1, cw, a0);
      axpy_skip(r1, sw, a1);
      axpy_skip(r0, 1.0, a2);
      if (k <= 2) { sm.top[k * n + l] = r0; sm.top[k * n + l + 1] = r1; }
      else { Pf[(size_t)k * n + l] = r0; Pf[(size_t)k * n + l + 1] = r1; }
    }
    L += 1;
    __syncthreads();
  }

  /* ---- reset, Robot.cpp:893-904 (dead entries are never read; downloads zero them) ---- */
  if (L > g.cap - g.headroom) {
    L = 0;
    if (tid == 0) st->resets += 1;
  }
  __syncthreads();
  const int nl_new = 3 + 2 * L;
  for (int i = tid; i < 3 * n; i += nt) { const int r = i / n, q = i % n; if (q < nl_new) Pf[(size_t)r * n + q] = sm.top[i]; }
  for (int j = tid; j < L; j += nt) {
    const int a = 3 + 2 * j;
    Pf[(size_t)a * n + a] = sm.diag[3 * j]; Pf[(size_t)a * n + a + 1] = sm.diag[3 * j + 1]; Pf[(size_t)(a + 1) * n + a + 1] = sm.diag[3 * j + 2];
  }
  for (int i = tid; i < n; i += nt) yf[i] = (i < nl_new) ? sm.y[i] : 0.0;
  if (tid == 0) { st->L = L; st->pose[0] = s_pose[0]; st->pose[1] = s_pose[1]; st->pose[2] = s_pose[2]; }
}

__global__ void k_batch_init(EkfBatchGeom g, double* Pg) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < g.B) {
    double* Pf = Pg + (size_t)f * g.n * g.n;
    Pf[0] = 0.05; Pf[(size_t)g.n + 1] = 0.05; Pf[(size_t)2 * g.n + 2] = 0.0;      /* Robot.cpp:27-30 */
  }
}

}  // namespace

struct ekf_batch {
  ekf_config cfg;
  EkfBatchGeom g;
  cudaStream_t stream;
  double* d_y; double* d_P; EkfBatchState* d_st;
  int max_m;
  double* d_in; double* h_in;       /* [u (3B) | z (2 m B) | R (4 m B)] */
  int* d_jout; int* h_jout;
  EkfBatchState* h_st;
  size_t smem_bytes;
  char err[256];
};

namespace {
#define CUB(call)                                                                             \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      snprintf(b->err, sizeof b->err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return EKF_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

size_t batch_smem(int n, int cap, int m) { return batch_smem_bytes(n, cap, m); }

int batch_ensure_m(ekf_batch* b, int m) {
  if (m <= b->max_m) return EKF_OK;
  CUB(cudaStreamSynchronize(b->stream));
  cudaFree(b->d_in); cudaFree(b->d_jout); cudaFreeHost(b->h_in); cudaFreeHost(b->h_jout);
  b->d_in = 0; b->d_jout = 0; b->h_in = 0; b->h_jout = 0;
  int cap = b->max_m > 0 ? b->max_m : 8;
  while (cap < m) cap *= 2;
  const size_t smem = batch_smem(b->g.n, b->g.cap, cap);
  if (smem > 227 * 1024) { snprintf(b->err, sizeof b->err, "scan of %d lines does not fit shared memory", m); return EKF_EINVAL; }
  b->max_m = cap;
  const size_t B = b->g.B;
  CUB(cudaMalloc(&b->d_in, (3 + 6 * (size_t)cap) * B * sizeof(double)));
  CUB(cudaMallocHost(&b->h_in, (3 + 6 * (size_t)cap) * B * sizeof(double)));
  CUB(cudaMalloc(&b->d_jout, (size_t)cap * B * sizeof(int)));
  CUB(cudaMallocHost(&b->h_jout, (size_t)cap * B * sizeof(int)));
  return EKF_OK;
}

int batch_launch(ekf_batch* b, const double* d_u, int m, const double* d_z, const double* d_R, int* d_jout) {
  const size_t smem = batch_smem(b->g.n, b->g.cap, m);
  if (smem > b->smem_bytes) {
    CUB(cudaFuncSetAttribute(k_batch_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b->smem_bytes = smem;
  }
  k_batch_scan<<<b->g.B, EKFB_THREADS, smem, b->stream>>>(b->g, b->d_y, b->d_P, b->d_st, d_u, d_z, d_R, m, d_jout);
  CUB(cudaGetLastError());
  return EKF_OK;
}
}  // namespace

extern "C" {

int ekf_batch_create(ekf_batch** out, const ekf_config* cfg, int n_filters) {
  if (!out || !cfg || n_filters < 1 || cfg->capacity_lines < 1) return EKF_EINVAL;
  ekf_batch* b = new ekf_batch();
  memset(b, 0, sizeof *b);
  b->cfg = *cfg;
  *out = b;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    snprintf(b->err, sizeof b->err, "no CUDA device: libekfcuda has no CPU fallback");
    return EKF_ECUDA;
  }
  CUB(cudaSetDevice(cfg->device));
  EkfBatchGeom& g = b->g;
  g.B = n_filters; g.cap = cfg->capacity_lines; g.n = 3 + 2 * g.cap; g.headroom = cfg->reset_headroom;
  g.gate = cfg->gate; g.enc_noise = cfg->encoder_noise;
  if (batch_smem(g.n, g.cap, 8) > 227 * 1024) {
    snprintf(b->err, sizeof b->err, "capacity %d does not fit shared memory; use ekf_create", g.cap);
    return EKF_EINVAL;
  }
  CUB(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  const size_t B = g.B, n = g.n;
  CUB(cudaMalloc(&b->d_y, B * n * sizeof(double)));
  CUB(cudaMalloc(&b->d_P, B * n * n * sizeof(double)));
  CUB(cudaMalloc(&b->d_st, B * sizeof(EkfBatchState)));
  CUB(cudaMallocHost(&b->h_st, B * sizeof(EkfBatchState)));
  CUB(cudaMemsetAsync(b->d_y, 0, B * n * sizeof(double), b->stream));
  CUB(cudaMemsetAsync(b->d_P, 0, B * n * n * sizeof(double), b->stream));
  CUB(cudaMemsetAsync(b->d_st, 0, B * sizeof(EkfBatchState), b->stream));
  k_batch_init<<<(g.B + 255) / 256, 256, 0, b->stream>>>(g, b->d_P);
  CUB(cudaGetLastError());
  int rc = batch_ensure_m(b, 8);
  if (rc) return rc;
  CUB(cudaStreamSynchronize(b->stream));
  return EKF_OK;
}

int ekf_batch_destroy(ekf_batch* b) {
  if (!b) return EKF_EINVAL;
  cudaSetDevice(b->cfg.device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  cudaFree(b->d_y); cudaFree(b->d_P); cudaFree(b->d_st); cudaFree(b->d_in); cudaFree(b->d_jout);
  cudaFreeHost(b->h_in); cudaFreeHost(b->h_jout); cudaFreeHost(b->h_st);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
  return EKF_OK;
}

const char* ekf_batch_last_error(const ekf_batch* b) { return b ? b->err : "null batch"; }

int ekf_batch_scan_device(ekf_batch* b, const double* d_u, int m, const double* d_z, const double* d_R, int* d_j_out) {
  if (!b || !d_u || m < 0 || (m > 0 && (!d_z || !d_R))) return EKF_EINVAL;
  CUB(cudaSetDevice(b->cfg.device));
  if (batch_smem(b->g.n, b->g.cap, m) > 227 * 1024) return EKF_EINVAL;
  return batch_launch(b, d_u, m, d_z, d_R, d_j_out);
}

int ekf_batch_scan(ekf_batch* b, const double* u, int m, const double* z, const double* R, int* j_out, double* pose) {
  if (!b || !u || m < 0 || (m > 0 && (!z || !R))) return EKF_EINVAL;
  CUB(cudaSetDevice(b->cfg.device));
  int rc = batch_ensure_m(b, m);
  if (rc) return rc;
  const size_t B = b->g.B;
  double* h = b->h_in;
  memcpy(h, u, 3 * B * sizeof(double));
  if (m > 0) {
    memcpy(h + 3 * B, z, 2 * (size_t)m * B * sizeof(double));
    memcpy(h + 3 * B + 2 * (size_t)m * B, R, 4 * (size_t)m * B * sizeof(double));
  }
  const size_t total = (3 + 6 * (size_t)m) * B;
  CUB(cudaMemcpyAsync(b->d_in, h, total * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  rc = batch_launch(b, b->d_in, m, b->d_in + 3 * B, b->d_in + 3 * B + 2 * (size_t)m * B, b->d_jout);
  if (rc) return rc;
  if (j_out && m > 0) CUB(cudaMemcpyAsync(b->h_jout, b->d_jout, (size_t)m * B * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CUB(cudaMemcpyAsync(b->h_st, b->d_st, B * sizeof(EkfBatchState), cudaMemcpyDeviceToHost, b->stream));
  CUB(cudaStreamSynchronize(b->stream));
  if (j_out && m > 0) memcpy(j_out, b->h_jout, (size_t)m * B * sizeof(int));
  int status = EKF_OK;
  for (size_t f = 0; f < B; ++f) {
    if (pose) memcpy(pose + 3 * f, b->h_st[f].pose, 3 * sizeof(double));
    if (b->h_st[f].sticky & EKF_STICKY_CAPACITY) status = EKF_ECAPACITY;
    else if ((b->h_st[f].sticky & EKF_STICKY_SINGULAR) && status == EKF_OK) status = EKF_ESINGULAR;
  }
  return status;
}

int ekf_batch_sync(ekf_batch* b) {
  if (!b) return EKF_EINVAL;
  CUB(cudaSetDevice(b->cfg.device));
  CUB(cudaStreamSynchronize(b->stream));
  return EKF_OK;
}

int ekf_batch_download(ekf_batch* b, int filter, double* y, double* P, int* n_lines, double pose[3]) {
  if (!b || filter < 0 || filter >= b->g.B) return EKF_EINVAL;
  CUB(cudaSetDevice(b->cfg.device));
  const size_t n = b->g.n;
  if (y) CUB(cudaMemcpyAsync(y, b->d_y + (size_t)filter * n, n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (P) CUB(cudaMemcpyAsync(P, b->d_P + (size_t)filter * n * n, n * n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CUB(cudaMemcpyAsync(b->h_st, b->d_st + filter, sizeof(EkfBatchState), cudaMemcpyDeviceToHost, b->stream));
  CUB(cudaStreamSynchronize(b->stream));
  if (n_lines) *n_lines = b->h_st[0].L;
  if (pose) memcpy(pose, b->h_st[0].pose, 3 * sizeof(double));
  /* the device keeps the upper triangle of the live part: mirror it and zero the rest (Robot::P_t0 layout) */
  const size_t nl = 3 + 2 * (size_t)b->h_st[0].L;
  if (P)
    for (size_t r = 0; r < n; ++r)
      for (size_t q = 0; q < n; ++q) {
        if (r >= nl || q >= nl) P[r * n + q] = 0.0;
        else if (q < r) P[r * n + q] = P[q * n + r];
      }
  if (y) for (size_t r = nl; r < n; ++r) y[r] = 0.0;
  return EKF_OK;
}

}  /* extern "C" */
