/* gsl_shim.h -- TEST INFRASTRUCTURE ONLY (oracle side).
 *
 * A from-scratch restatement of the small GSL subset that the reference's
 * slam_ros/Robot.cpp, lineFitting.cpp and simplifyPath.cpp call, so that the
 * UNMODIFIED reference translation units can be compiled in an image that has
 * no libgsl (GSL is a system, un-vendored, un-pinned
 * dependency of the reference -- slam_ros/CMakeLists.txt:44-51 links "gsl" and
 * "gslcblas" by bare name; ROS Kinetic => Ubuntu 16.04 => GSL 2.1 by inference).
 *
 * What is restated is GSL's *published* reference algorithm, not its source:
 *   - gsl_blas_dgemm -> row-major reference CBLAS dgemm: C is first scaled by
 *     beta (beta == 0 stores 0.0); NN / TN run k-outer, i-middle, j-inner with
 *     the "skip when alpha*A(i,k) == 0" short cut; NT / TT run i,j-outer with a
 *     scalar dot product accumulated from 0.0 in k order.  No FMA (build with
 *     -ffp-contract=off), no blocking.
 *   - gsl_linalg_LU_decomp: unblocked Gaussian elimination with partial
 *     pivoting, first-largest pivot, sub-column scaled by division (GSL <= 2.5).
 *   - gsl_linalg_LU_invert: identity, then per column permute / unit-lower
 *     forward substitution / upper back substitution (reference dtrsv order).
 *   - gsl_matrix element-wise helpers return GSL_EBADLEN on shape mismatch and
 *     gsl_matrix_get/set are range-checked: with the error handler off an
 *     out-of-range get returns 0 and a set is dropped (Robot.cpp's debug loops
 *     at :171-176, :196-201, :634-638 depend on that -- SURVEY.md Q13).
 *   - gsl_eigen_nonsymmv is only used by Robot::getEllipse on a 2x2; the shim
 *     provides a closed-form real 2x2 eigen-decomposition with unit-norm
 *     eigenvectors.  Eigenvector SIGN is not pinned to GSL's.
 *
 * Nothing in the shipped library includes this file.
 */
#ifndef EKF_ORACLE_GSL_SHIM_H
#define EKF_ORACLE_GSL_SHIM_H

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <random>   /* slam_ros/main.cpp:44 relies on this arriving transitively */

#define GSL_SUCCESS 0
#define GSL_FAILURE (-1)
#define GSL_EDOM 1
#define GSL_EINVAL 4
#define GSL_EBADLEN 19
#define GSL_ENOTSQR 20

extern "C++" {

typedef void gsl_error_handler_t(const char*, const char*, int, int);

struct gsl_block { size_t size; double* data; };
struct gsl_matrix { size_t size1, size2, tda; double* data; gsl_block* block; int owner; };
struct gsl_matrix_view { gsl_matrix matrix; };
typedef gsl_matrix_view _gsl_matrix_view;
struct gsl_vector { size_t size, stride; double* data; gsl_block* block; int owner; };
struct gsl_vector_view { gsl_vector vector; };
struct gsl_permutation { size_t size; size_t* data; };

struct gsl_complex { double dat[2]; };
#define GSL_REAL(z) ((z).dat[0])
#define GSL_IMAG(z) ((z).dat[1])
struct gsl_vector_complex { size_t size, stride; double* data; void* block; int owner; };
struct gsl_vector_complex_view { gsl_vector_complex vector; };
struct gsl_matrix_complex { size_t size1, size2, tda; double* data; void* block; int owner; };
struct gsl_eigen_nonsymmv_workspace { size_t size; };
typedef enum { GSL_EIGEN_SORT_VAL_ASC, GSL_EIGEN_SORT_VAL_DESC,
               GSL_EIGEN_SORT_ABS_ASC, GSL_EIGEN_SORT_ABS_DESC } gsl_eigen_sort_t;

struct gsl_function { double (*function)(double, void*); void* params; };

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;
typedef CBLAS_TRANSPOSE_t CBLAS_TRANSPOSE;

/* counters the harness can read: how many range errors the literal path made */
struct gsl_shim_counters { unsigned long range_errors; unsigned long badlen_errors; };
inline gsl_shim_counters& gsl_shim_stats() { static gsl_shim_counters c = {0, 0}; return c; }

inline gsl_error_handler_t* gsl_set_error_handler_off() { return 0; }
inline gsl_error_handler_t* gsl_set_error_handler(gsl_error_handler_t*) { return 0; }

inline const char* gsl_strerror(int e) {
  switch (e) {
    case GSL_SUCCESS: return "success";
    case GSL_FAILURE: return "failure";
    case GSL_EDOM: return "input domain error";
    case GSL_EINVAL: return "invalid argument supplied by user";
    case GSL_EBADLEN: return "matrix/vector sizes are not conformant";
    case GSL_ENOTSQR: return "matrix not square";
    default: return "unknown error code";
  }
}

/* ---- matrices ---------------------------------------------------------- */
inline gsl_matrix* gsl_matrix_alloc(size_t n1, size_t n2) {
  gsl_matrix* m = (gsl_matrix*)std::malloc(sizeof(gsl_matrix));
  gsl_block* b = (gsl_block*)std::malloc(sizeof(gsl_block));
  b->size = n1 * n2;
  b->data = (double*)std::malloc(sizeof(double) * (n1 * n2 ? n1 * n2 : 1));
#ifdef GSL_SHIM_ZERO_ALLOC   /* one legal instance of "indeterminate": used for the line-extraction reference build */
  std::memset(b->data, 0, sizeof(double) * (n1 * n2 ? n1 * n2 : 1));
#endif
  m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = b->data; m->block = b; m->owner = 1;
  return m;
}
inline gsl_matrix* gsl_matrix_calloc(size_t n1, size_t n2) {
  gsl_matrix* m = gsl_matrix_alloc(n1, n2);
  std::memset(m->data, 0, sizeof(double) * n1 * n2);
  return m;
}
inline void gsl_matrix_free(gsl_matrix* m) {
  if (!m) return;
  if (m->owner && m->block) { std::free(m->block->data); std::free(m->block); }
  std::free(m);
}
inline gsl_matrix_view gsl_matrix_view_array(double* base, size_t n1, size_t n2) {
  gsl_matrix_view v;
  v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = n2;
  v.matrix.data = base; v.matrix.block = 0; v.matrix.owner = 0;
  return v;
}
inline gsl_matrix_view gsl_matrix_submatrix(gsl_matrix* m, size_t i, size_t j, size_t n1, size_t n2) {
  gsl_matrix_view v;
  v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = m->tda;
  v.matrix.data = m->data + i * m->tda + j; v.matrix.block = m->block; v.matrix.owner = 0;
  return v;
}
inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) {
  if (i >= m->size1 || j >= m->size2) { gsl_shim_stats().range_errors++; return 0; }
  return m->data[i * m->tda + j];
}
inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) {
  if (i >= m->size1 || j >= m->size2) { gsl_shim_stats().range_errors++; return; }
  m->data[i * m->tda + j] = x;
}
inline void gsl_matrix_set_all(gsl_matrix* m, double x) {
  for (size_t i = 0; i < m->size1; ++i)
    for (size_t j = 0; j < m->size2; ++j) m->data[i * m->tda + j] = x;
}
inline void gsl_matrix_set_zero(gsl_matrix* m) { gsl_matrix_set_all(m, 0.0); }
inline void gsl_matrix_set_identity(gsl_matrix* m) {
  for (size_t i = 0; i < m->size1; ++i)
    for (size_t j = 0; j < m->size2; ++j) m->data[i * m->tda + j] = (i == j) ? 1.0 : 0.0;
}
inline int gsl_matrix_memcpy(gsl_matrix* d, const gsl_matrix* s) {
  if (d->size1 != s->size1 || d->size2 != s->size2) { gsl_shim_stats().badlen_errors++; return GSL_EBADLEN; }
  for (size_t i = 0; i < s->size1; ++i)
    for (size_t j = 0; j < s->size2; ++j) d->data[i * d->tda + j] = s->data[i * s->tda + j];
  return GSL_SUCCESS;
}
inline int gsl_matrix_transpose_memcpy(gsl_matrix* d, const gsl_matrix* s) {
  if (d->size2 != s->size1 || d->size1 != s->size2) { gsl_shim_stats().badlen_errors++; return GSL_EBADLEN; }
  for (size_t i = 0; i < d->size1; ++i)
    for (size_t j = 0; j < d->size2; ++j) d->data[i * d->tda + j] = s->data[j * s->tda + i];
  return GSL_SUCCESS;
}
inline int gsl_matrix_add(gsl_matrix* a, const gsl_matrix* b) {
  if (a->size1 != b->size1 || a->size2 != b->size2) { gsl_shim_stats().badlen_errors++; return GSL_EBADLEN; }
  for (size_t i = 0; i < a->size1; ++i)
    for (size_t j = 0; j < a->size2; ++j) a->data[i * a->tda + j] += b->data[i * b->tda + j];
  return GSL_SUCCESS;
}
inline int gsl_matrix_sub(gsl_matrix* a, const gsl_matrix* b) {
  if (a->size1 != b->size1 || a->size2 != b->size2) { gsl_shim_stats().badlen_errors++; return GSL_EBADLEN; }
  for (size_t i = 0; i < a->size1; ++i)
    for (size_t j = 0; j < a->size2; ++j) a->data[i * a->tda + j] -= b->data[i * b->tda + j];
  return GSL_SUCCESS;
}
inline int gsl_matrix_swap_rows(gsl_matrix* m, size_t i, size_t j) {
  if (i >= m->size1 || j >= m->size1) return GSL_EINVAL;
  if (i != j)
    for (size_t k = 0; k < m->size2; ++k) {
      double t = m->data[i * m->tda + k];
      m->data[i * m->tda + k] = m->data[j * m->tda + k];
      m->data[j * m->tda + k] = t;
    }
  return GSL_SUCCESS;
}

/* ---- permutations ------------------------------------------------------ */
inline gsl_permutation* gsl_permutation_alloc(size_t n) {
  gsl_permutation* p = (gsl_permutation*)std::malloc(sizeof(gsl_permutation));
  p->size = n; p->data = (size_t*)std::malloc(sizeof(size_t) * (n ? n : 1));
  return p;
}
inline void gsl_permutation_free(gsl_permutation* p) { if (p) { std::free(p->data); std::free(p); } }
inline void gsl_permutation_init(gsl_permutation* p) { for (size_t i = 0; i < p->size; ++i) p->data[i] = i; }

/* ---- reference row-major dgemm ----------------------------------------- */
inline int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TA, CBLAS_TRANSPOSE_t TB, double alpha,
                          const gsl_matrix* A, const gsl_matrix* B, double beta, gsl_matrix* C) {
  const size_t M = C->size1, N = C->size2;
  const size_t MA = (TA == CblasNoTrans) ? A->size1 : A->size2;
  const size_t NA = (TA == CblasNoTrans) ? A->size2 : A->size1;
  const size_t MB = (TB == CblasNoTrans) ? B->size1 : B->size2;
  const size_t NB = (TB == CblasNoTrans) ? B->size2 : B->size1;
  if (!(M == MA && N == NB && NA == MB)) { gsl_shim_stats().badlen_errors++; return GSL_EBADLEN; }
  const size_t K = NA;
  const double* a = A->data; const double* b = B->data; double* c = C->data;
  const size_t lda = A->tda, ldb = B->tda, ldc = C->tda;
  if (alpha == 0.0 && beta == 1.0) return GSL_SUCCESS;
  if (beta == 0.0) {
    for (size_t i = 0; i < M; ++i) for (size_t j = 0; j < N; ++j) c[ldc * i + j] = 0.0;
  } else if (beta != 1.0) {
    for (size_t i = 0; i < M; ++i) for (size_t j = 0; j < N; ++j) c[ldc * i + j] *= beta;
  }
  if (alpha == 0.0) return GSL_SUCCESS;
  const bool ta = (TA != CblasNoTrans), tb = (TB != CblasNoTrans);
  if (!tb) {                       /* NN and TN: outer-product order with zero skip */
    for (size_t k = 0; k < K; ++k)
      for (size_t i = 0; i < M; ++i) {
        const double t = alpha * (ta ? a[lda * k + i] : a[lda * i + k]);
        if (t != 0.0)
          for (size_t j = 0; j < N; ++j) c[ldc * i + j] += t * b[ldb * k + j];
      }
  } else {                         /* NT and TT: dot-product order */
    for (size_t i = 0; i < M; ++i)
      for (size_t j = 0; j < N; ++j) {
        double t = 0.0;
        for (size_t k = 0; k < K; ++k) t += (ta ? a[lda * k + i] : a[lda * i + k]) * b[ldb * j + k];
        c[ldc * i + j] += alpha * t;
      }
  }
  return GSL_SUCCESS;
}

/* ---- LU (unblocked, partial pivoting) ----------------------------------- */
inline int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum) {
  if (A->size1 != A->size2) return GSL_ENOTSQR;
  if (p->size != A->size1) return GSL_EBADLEN;
  const size_t N = A->size1;
  *signum = 1;
  gsl_permutation_init(p);
  for (size_t j = 0; j + 1 < N; ++j) {
    double big = std::fabs(A->data[j * A->tda + j]);
    size_t piv = j;
    for (size_t i = j + 1; i < N; ++i) {
      const double v = std::fabs(A->data[i * A->tda + j]);
      if (v > big) { big = v; piv = i; }
    }
    if (piv != j) {
      gsl_matrix_swap_rows(A, j, piv);
      size_t t = p->data[j]; p->data[j] = p->data[piv]; p->data[piv] = t;
      *signum = -*signum;
    }
    const double ajj = A->data[j * A->tda + j];
    if (ajj != 0.0)
      for (size_t i = j + 1; i < N; ++i) {
        const double l = A->data[i * A->tda + j] / ajj;
        A->data[i * A->tda + j] = l;
        for (size_t k = j + 1; k < N; ++k)
          A->data[i * A->tda + k] = A->data[i * A->tda + k] - l * A->data[j * A->tda + k];
      }
  }
  return GSL_SUCCESS;
}
inline int gsl_linalg_LU_invert(const gsl_matrix* LU, const gsl_permutation* p, gsl_matrix* inv) {
  const size_t N = LU->size1;
  for (size_t i = 0; i < N; ++i)
    if (LU->data[i * LU->tda + i] == 0.0) return GSL_EDOM;      /* singular */
  gsl_matrix_set_identity(inv);
  double* x = (double*)std::malloc(sizeof(double) * (N ? N : 1));
  double* t = (double*)std::malloc(sizeof(double) * (N ? N : 1));
  for (size_t col = 0; col < N; ++col) {
    for (size_t i = 0; i < N; ++i) t[i] = inv->data[i * inv->tda + col];
    for (size_t i = 0; i < N; ++i) x[i] = t[p->data[i]];         /* apply the row permutation */
    for (size_t i = 1; i < N; ++i) {                              /* unit lower, forward */
      double s = x[i];
      for (size_t j = 0; j < i; ++j) s -= LU->data[i * LU->tda + j] * x[j];
      x[i] = s;
    }
    x[N - 1] = x[N - 1] / LU->data[(N - 1) * LU->tda + (N - 1)]; /* upper, backward */
    for (size_t i = N - 1; i-- > 0;) {
      double s = x[i];
      for (size_t j = i + 1; j < N; ++j) s -= LU->data[i * LU->tda + j] * x[j];
      x[i] = s / LU->data[i * LU->tda + i];
    }
    for (size_t i = 0; i < N; ++i) inv->data[i * inv->tda + col] = x[i];
  }
  std::free(x); std::free(t);
  return GSL_SUCCESS;
}

/* ---- complex containers + 2x2 eigen (getEllipse only) -------------------- */
inline gsl_vector_complex* gsl_vector_complex_alloc(size_t n) {
  gsl_vector_complex* v = (gsl_vector_complex*)std::malloc(sizeof(gsl_vector_complex));
  v->size = n; v->stride = 1; v->data = (double*)std::calloc(2 * (n ? n : 1), sizeof(double));
  v->block = 0; v->owner = 1;
  return v;
}
inline void gsl_vector_complex_free(gsl_vector_complex* v) { if (v) { if (v->owner) std::free(v->data); std::free(v); } }
inline gsl_complex gsl_vector_complex_get(const gsl_vector_complex* v, size_t i) {
  gsl_complex z; z.dat[0] = v->data[2 * i * v->stride]; z.dat[1] = v->data[2 * i * v->stride + 1]; return z;
}
inline gsl_matrix_complex* gsl_matrix_complex_alloc(size_t n1, size_t n2) {
  gsl_matrix_complex* m = (gsl_matrix_complex*)std::malloc(sizeof(gsl_matrix_complex));
  m->size1 = n1; m->size2 = n2; m->tda = n2;
  m->data = (double*)std::calloc(2 * (n1 * n2 ? n1 * n2 : 1), sizeof(double)); m->block = 0; m->owner = 1;
  return m;
}
inline void gsl_matrix_complex_free(gsl_matrix_complex* m) { if (m) { if (m->owner) std::free(m->data); std::free(m); } }
inline gsl_vector_complex_view gsl_matrix_complex_column(gsl_matrix_complex* m, size_t j) {
  gsl_vector_complex_view v;
  v.vector.size = m->size1; v.vector.stride = m->tda; v.vector.data = m->data + 2 * j; v.vector.block = 0; v.vector.owner = 0;
  return v;
}
inline gsl_eigen_nonsymmv_workspace* gsl_eigen_nonsymmv_alloc(size_t n) {
  gsl_eigen_nonsymmv_workspace* w = (gsl_eigen_nonsymmv_workspace*)std::malloc(sizeof(gsl_eigen_nonsymmv_workspace));
  w->size = n; return w;
}
inline void gsl_eigen_nonsymmv_free(gsl_eigen_nonsymmv_workspace* w) { std::free(w); }
/* closed-form real 2x2; complex pairs are reported with zero eigenvectors (never hit for a covariance) */
inline int gsl_eigen_nonsymmv(gsl_matrix* A, gsl_vector_complex* eval, gsl_matrix_complex* evec,
                              gsl_eigen_nonsymmv_workspace*) {
  if (A->size1 != 2 || A->size2 != 2) return GSL_EINVAL;
  const double a = A->data[0], b = A->data[1], c = A->data[A->tda], d = A->data[A->tda + 1];
  const double tr = a + d, half = 0.5 * (a - d), disc = half * half + b * c;
  if (disc < 0.0) {
    const double im = std::sqrt(-disc);
    eval->data[0] = 0.5 * tr; eval->data[1] = im; eval->data[2] = 0.5 * tr; eval->data[3] = -im;
    std::memset(evec->data, 0, sizeof(double) * 8);
    return GSL_SUCCESS;
  }
  const double rt = std::sqrt(disc);
  const double l[2] = {0.5 * tr + rt, 0.5 * tr - rt};
  for (int k = 0; k < 2; ++k) {
    double vx, vy;
    if (b != 0.0)      { vx = b;          vy = l[k] - a; }
    else if (c != 0.0) { vx = l[k] - d;   vy = c; }
    else               { vx = (k == 0) == (a >= d) ? 1.0 : 0.0; vy = 1.0 - vx; }
    const double nrm = std::sqrt(vx * vx + vy * vy);
    if (nrm > 0.0) { vx /= nrm; vy /= nrm; }
    eval->data[2 * k] = l[k]; eval->data[2 * k + 1] = 0.0;
    evec->data[2 * (0 * evec->tda + k)] = vx; evec->data[2 * (0 * evec->tda + k) + 1] = 0.0;
    evec->data[2 * (1 * evec->tda + k)] = vy; evec->data[2 * (1 * evec->tda + k) + 1] = 0.0;
  }
  return GSL_SUCCESS;
}
inline int gsl_eigen_nonsymmv_sort(gsl_vector_complex* eval, gsl_matrix_complex* evec, gsl_eigen_sort_t how) {
  if (eval->size != 2) return GSL_EINVAL;
  const double m0 = std::hypot(eval->data[0], eval->data[1]), m1 = std::hypot(eval->data[2], eval->data[3]);
  bool swap = false;
  switch (how) {
    case GSL_EIGEN_SORT_ABS_ASC:  swap = m1 < m0; break;
    case GSL_EIGEN_SORT_ABS_DESC: swap = m1 > m0; break;
    case GSL_EIGEN_SORT_VAL_ASC:  swap = eval->data[2] < eval->data[0]; break;
    case GSL_EIGEN_SORT_VAL_DESC: swap = eval->data[2] > eval->data[0]; break;
  }
  if (swap) {
    for (int q = 0; q < 2; ++q) { double t = eval->data[q]; eval->data[q] = eval->data[2 + q]; eval->data[2 + q] = t; }
    for (size_t r = 0; r < 2; ++r)
      for (int q = 0; q < 2; ++q) {
        double* x = &evec->data[2 * (r * evec->tda + 0) + q];
        double* y = &evec->data[2 * (r * evec->tda + 1) + q];
        double t = *x; *x = *y; *y = t;
      }
  }
  return GSL_SUCCESS;
}

/* ---- declared because the headers are #included; only referenced from
 *      commented-out code in lineFitting.cpp:312-355 ----------------------- */
inline int gsl_deriv_forward(const gsl_function* f, double x, double h, double* result, double* abserr) {
  const double f0 = f->function(x, f->params), f1 = f->function(x + h, f->params);
  *result = (f1 - f0) / h; *abserr = std::fabs(*result) * 1e-8; return GSL_SUCCESS;
}
inline double gsl_sf_bessel_J0(double x) { return ::j0(x); }

} /* extern "C++" */
#endif
