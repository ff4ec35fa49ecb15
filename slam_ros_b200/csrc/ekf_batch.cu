/* ekf_batch.cu -- independent filter instances (Monte-Carlo batches, BASELINE.json configs[3]).
 *
 * One thread block per filter.  The whole Robot::localize (slam_ros/Robot.cpp:126-943) of that filter
 * runs inside one kernel launch with its covariance staged in shared memory: P is read from HBM once
 * and written back once per scan (16 n^2 bytes), however many lines match.  Because shared memory is
 * cheap at this size (n = 103 -> 85 KB) the block keeps the FULL, non-symmetrised n x n matrix and
 * applies the reference's full-matrix update (Robot.cpp:564-568) element for element, so apart from
 * CUDA's sin/cos the arithmetic is the reference's own.  Filters never communicate; a multi-GPU batch
 * is N independent ekf_batch objects, one per device.
 */
#include "../../include/ekf.h"
#include "ekf_internal.h"
#include "ekf_device.cuh"

#include <stdio.h>
#include <string.h>

#define EKFB_THREADS 256

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

/* Robot::localize for filter blockIdx.x */
__global__ void __launch_bounds__(EKFB_THREADS) k_batch_scan(EkfBatchGeom g, double* __restrict__ Yg,
                                                             double* __restrict__ Pg, EkfBatchState* __restrict__ Sg,
                                                             const double* __restrict__ U, const double* __restrict__ Z,
                                                             const double* __restrict__ Rm, int m, int* __restrict__ Jout) {
  extern __shared__ double smem[];
  const int n = g.n, tid = threadIdx.x, nt = blockDim.x;
  const int f = blockIdx.x;
  double* Ps = smem;                          /* n x n */
  double* ys = Ps + (size_t)n * n;            /* n */
  double2* Ks = reinterpret_cast<double2*>(ys + n + ((n * n + n) & 1));   /* n  (16-byte aligned) */
  double2* KSs = Ks + n;                      /* n */
  int* ext = reinterpret_cast<int*>(KSs + n); /* m */
  unsigned char* matched = reinterpret_cast<unsigned char*>(ext + m);   /* cap */
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

  for (int i = tid; i < n * n; i += nt) Ps[i] = Pf[i];
  for (int i = tid; i < n; i += nt) ys[i] = yf[i];
  for (int i = tid; i < g.cap; i += nt) matched[i] = 0;
  if (tid == 0) {
    st->sticky = 0;                           /* status reports this scan only */
    s_L = st->L; s_ne = 0; s_nmatch = 0;
    s_pose[0] = st->pose[0]; s_pose[1] = st->pose[1]; s_pose[2] = st->pose[2];
  }
  __syncthreads();

  /* ---- prediction, Robot.cpp:130-258 ---- */
  const double u0 = u[0], u2 = u[2];
  const double ang = add_rn(s_pose[2], __ddiv_rn(u2, 2.0));
  const double ca = cos(ang), sa = sin(ang);
  const double F02 = mul_rn(-u0, sa), F12 = mul_rn(u0, ca);
  int L = s_L;
  int nl = 3 + 2 * L;
  for (int j = tid; j < nl; j += nt) {                                /* :242 */
    const double p0 = Ps[j], p1 = Ps[n + j], p2 = Ps[2 * n + j];
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    axpy_skip(t0, 1.0, p0); axpy_skip(t1, 1.0, p1);
    axpy_skip(t0, F02, p2); axpy_skip(t1, F12, p2); axpy_skip(t2, 1.0, p2);
    Ps[j] = t0; Ps[n + j] = t1; Ps[2 * n + j] = t2;
  }
  __syncthreads();
  for (int i = tid; i < nl; i += nt) {                                /* :246 */
    double* Ti = Ps + (size_t)i * n;
    double c0 = 0.0; c0 = add_rn(c0, mul_rn(Ti[0], 1.0)); c0 = add_rn(c0, mul_rn(Ti[2], F02));
    double c1 = 0.0; c1 = add_rn(c1, mul_rn(Ti[1], 1.0)); c1 = add_rn(c1, mul_rn(Ti[2], F12));
    Ti[0] = add_rn(0.0, c0); Ti[1] = add_rn(0.0, c1);
  }
  __syncthreads();
  if (tid == 0) {
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
      for (int j = 0; j < 3; ++j) {
        double t = 0.0;
        for (int k = 0; k < 3; ++k) t = add_rn(t, mul_rn(FQ[i][k], Fu[j][k]));
        Ps[i * n + j] = add_rn(Ps[i * n + j], add_rn(0.0, mul_rn(1.0, t)));
      }
    s_xpre[0] = add_rn(s_pose[0], mul_rn(u0, ca));
    s_xpre[1] = add_rn(s_pose[1], mul_rn(u0, sa));
    s_xpre[2] = add_rn(s_pose[2], u2);
  }
  __syncthreads();

  /* ---- the observed lines in order, Robot.cpp:298-645 ---- */
  for (int i = 0; i < m; ++i) {
    const double z0 = z[2 * i], z1 = z[2 * i + 1];
    const double Rl[4] = {R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]};
    const double xp[3] = {s_xpre[0], s_xpre[1], s_xpre[2]};
    int cand = EKF_NO_MATCH;
    for (int j = tid; j < L; j += nt) {
      if (matched[j] || j >= cand) continue;
      const int idx[5] = {0, 1, 2, 3 + 2 * j, 4 + 2 * j};
      double Cm[5][5];
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int c = 0; c < 5; ++c) Cm[a][c] = Ps[(size_t)idx[a] * n + idx[c]];
      Gate G;
      gate_from_block(Cm, ys[idx[3]], ys[idx[4]], xp, z0, z1, Rl, G);
      if (G.singular) atomicOr(&st->sticky, EKF_STICKY_SINGULAR);
      else if (!(sqrt(fabs(G.d2)) > g.gate)) cand = j;
    }
    cand = __reduce_min_sync(0xffffffffu, cand);
    if ((tid & 31) == 0) s_min[tid >> 5] = cand;
    __syncthreads();
    if (tid < 32) {
      int v = (tid < EKFB_THREADS / 32) ? s_min[tid] : EKF_NO_MATCH;
      v = __reduce_min_sync(0xffffffffu, v);
      if (tid == 0) {
        s_best = v;
        if (v == EKF_NO_MATCH) { ext[s_ne++] = i; if (jout) jout[i] = -1; }
        else {
          const int idx[5] = {0, 1, 2, 3 + 2 * v, 4 + 2 * v};
          double Cm[5][5];
          for (int a = 0; a < 5; ++a)
            for (int c = 0; c < 5; ++c) Cm[a][c] = Ps[(size_t)idx[a] * n + idx[c]];
          gate_from_block(Cm, ys[idx[3]], ys[idx[4]], xp, z0, z1, Rl, sG);
        }
      }
    }
    __syncthreads();
    const int jb = s_best;
    if (jb == EKF_NO_MATCH) continue;
    const int a = 3 + 2 * jb, bb = a + 1;
    for (int r = tid; r < nl; r += nt) {                              /* :516-560 */
      const double* Pr = Ps + (size_t)r * n;
      gain_row(sG, Pr[0], Pr[1], Pr[2], Pr[a], Pr[bb], Ks[r], KSs[r]);
    }
    __syncthreads();
    {                                                                 /* :564-568 full matrix */
      const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
      for (int r = warp; r < nl; r += nw) {
        const double2 ks = KSs[r];
        double* Pr = Ps + (size_t)r * n;
        for (int q = lane; q < nl; q += 32) Pr[q] = sub_rank2(Pr[q], ks, Ks[q]);
      }
    }
    for (int r = 3 + tid; r < nl; r += nt) {                          /* :585-589 */
      double t = 0.0;
      axpy_skip(t, Ks[r].x, sG.v[0]); axpy_skip(t, Ks[r].y, sG.v[1]);
      ys[r] = add_rn(ys[r], t);
    }
    if (tid == 0) {
      double yn[3];
      for (int r = 0; r < 3; ++r) {
        double t = 0.0;
        axpy_skip(t, Ks[r].x, sG.v[0]); axpy_skip(t, Ks[r].y, sG.v[1]);
        yn[r] = add_rn(s_xpre[r], t);
      }
      normalize_radian(yn[2]);
      for (int r = 0; r < 3; ++r) { ys[r] = yn[r]; s_pose[r] = yn[r]; s_xpre[r] = yn[r]; }
      matched[jb] = 1; s_nmatch++;
      if (jout) jout[i] = jb;
    }
    __syncthreads();
  }

  /* ---- Robot.cpp:702-716 ---- */
  if (tid == 0) {
    if (m == 0 || s_nmatch == 0) {
      ys[0] = s_xpre[0]; ys[1] = s_xpre[1]; ys[2] = s_xpre[2];
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
        const int i = ext[e];
        double alfa = z[2 * i], r = z[2 * i + 1];
        const double Rl[4] = {R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]};
        r = add_rn(r, add_rn(mul_rn(s_pose[0], cos(alfa)), mul_rn(s_pose[1], sin(alfa))));
        alfa = add_rn(alfa, s_pose[2]);
        const double cw = cos(alfa), sw = sin(alfa);
        const double Gx[2][3] = {{0.0, 0.0, 1.0}, {cw, sw, 0.0}};
        const double Gl[2][2] = {{1.0, 0.0}, {sub_rn(mul_rn(ys[1], cw), mul_rn(ys[0], sw)), 1.0}};
        normalize_radian(alfa);
        ys[l] = alfa; ys[l + 1] = r;
        s_cs[0] = cw; s_cs[1] = sw;
        double GP[2][3] = {{0, 0, 0}, {0, 0, 0}};
        for (int k = 0; k < 3; ++k)
          for (int ii = 0; ii < 2; ++ii) {
            const double t = mul_rn(1.0, Gx[ii][k]);
            if (t != 0.0) for (int jj = 0; jj < 3; ++jj) GP[ii][jj] = add_rn(GP[ii][jj], mul_rn(t, Ps[k * n + jj]));
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
        Ps[(size_t)l * n + l] = Pll[0][0]; Ps[(size_t)l * n + l + 1] = Pll[0][1];
        Ps[(size_t)(l + 1) * n + l] = Pll[1][0]; Ps[(size_t)(l + 1) * n + l + 1] = Pll[1][1];
      }
    }
    __syncthreads();
    if (s_stop) break;
    const double cw = s_cs[0], sw = s_cs[1];
    for (int j = tid; j < l; j += nt) {                               /* :856-860 */
      double r0 = 0.0, r1 = 0.0;
      axpy_skip(r1, cw, Ps[j]);
      axpy_skip(r1, sw, Ps[n + j]);
      axpy_skip(r0, 1.0, Ps[2 * n + j]);
      Ps[(size_t)l * n + j] = r0; Ps[(size_t)(l + 1) * n + j] = r1;
      Ps[(size_t)j * n + l] = r0; Ps[(size_t)j * n + l + 1] = r1;
    }
    L += 1;
    __syncthreads();
  }

  /* ---- reset, Robot.cpp:893-904 ---- */
  if (L > g.cap - g.headroom) {
    for (int i = 3 + tid; i < n; i += nt) ys[i] = 0.0;
    for (int i = tid; i < n * n; i += nt) { const int r = i / n, q = i % n; if (r >= 3 || q >= 3) Ps[i] = 0.0; }
    L = 0;
    if (tid == 0) st->resets += 1;
  }
  __syncthreads();
  for (int i = tid; i < n * n; i += nt) Pf[i] = Ps[i];
  for (int i = tid; i < n; i += nt) yf[i] = ys[i];
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

size_t batch_smem(int n, int cap, int m) {
  size_t bytes = ((size_t)n * n + n + (((size_t)n * n + n) & 1)) * sizeof(double) + 2 * (size_t)n * sizeof(double2) +
                 (size_t)m * sizeof(int) + (size_t)cap;
  return (bytes + 15) & ~(size_t)15;
}

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
  return EKF_OK;
}

}  /* extern "C" */
