// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Small fixed-size linear algebra used by the CPU restatement of the reference's
// hot path.  The reference delegates these to Eigen (un-vendored, version unpinned:
// SURVEY.md section 8c / Appendix B), so each routine restates the *published*
// Eigen 3.4 algorithm it stands in for.  Summation order inside Eigen's small
// products is not recoverable from /root/reference; we use plain left-to-right
// order everywhere ("parity unpinned" at the 1-ulp level).
//
//   jacobi_svd<N>        <- Eigen::JacobiSVD<Matrix<double,N,N>> (two-sided Jacobi, square => no QR
//                           preconditioner), used at NDT:127-129 (6x6 solve) and FG:273 (3x3).
//   self_adjoint_eigen3  <- Eigen::SelfAdjointEigenSolver<Matrix3d>::compute (tridiagonalise + implicit
//                           symmetric QR), used at VGC:333-335.
//   inverse3             <- Eigen::Matrix3d::inverse() (cofactors), VGC:355,359, FG:149.
//   ldlt_solve6          <- Eigen::LDLT<Matrix<double,6,6>> (pivoted, lower), LSQ:111,136.
//   euler_angles_012     <- Matrix3f::eulerAngles(0,1,2), NDT:109.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>

namespace lgs_oracle {

// ----------------------------------------------------------------------------------------------
// Jacobi rotations (Eigen/src/Jacobi/Jacobi.h semantics)
template <typename T>
struct RotT {
  T c, s;
};
using Rot = RotT<double>;
template <typename T>
inline RotT<T> rot_mul(const RotT<T>& a, const RotT<T>& b) { return {a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c}; }
template <typename T>
inline RotT<T> rot_T(const RotT<T>& a) { return {a.c, -a.s}; }

// rows p,q of an NxN row-major matrix:  x' = c x + s y ; y' = -s x + c y
template <int N, typename T>
inline void apply_left(T* m, int p, int q, const RotT<T>& j) {
  for (int i = 0; i < N; i++) {
    T x = m[p * N + i], y = m[q * N + i];
    m[p * N + i] = j.c * x + j.s * y;
    m[q * N + i] = -j.s * x + j.c * y;
  }
}
// columns p,q:  x' = c x - s y ; y' = s x + c y
template <int N, typename T>
inline void apply_right(T* m, int p, int q, const RotT<T>& j) {
  for (int i = 0; i < N; i++) {
    T x = m[i * N + p], y = m[i * N + q];
    m[i * N + p] = j.c * x - j.s * y;
    m[i * N + q] = j.s * x + j.c * y;
  }
}

template <typename T>
inline bool make_jacobi(T x, T y, T z, RotT<T>* r) {
  T deno = T(2) * std::fabs(y);
  if (deno < std::numeric_limits<T>::min()) {
    r->c = T(1);
    r->s = T(0);
    return false;
  }
  T tau = (x - z) / deno;
  T w = std::sqrt(tau * tau + T(1));
  T t = (tau > T(0)) ? T(1) / (tau + w) : T(1) / (tau - w);
  T sign_t = t > T(0) ? T(1) : T(-1);
  T n = T(1) / std::sqrt(t * t + T(1));
  r->s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  r->c = n;
  return true;
}

template <typename T>
inline void real_2x2_jacobi_svd(T mpp, T mpq, T mqp, T mqq, RotT<T>* j_left, RotT<T>* j_right) {
  T m[4] = {mpp, mpq, mqp, mqq};
  RotT<T> rot1;
  T t = m[0] + m[3];
  T d = m[2] - m[1];
  if (std::fabs(d) < std::numeric_limits<T>::min()) {
    rot1.s = T(0);
    rot1.c = T(1);
  } else {
    T u = t / d;
    T tmp = std::sqrt(T(1) + u * u);
    rot1.s = T(1) / tmp;
    rot1.c = u / tmp;
  }
  apply_left<2, T>(m, 0, 1, rot1);
  make_jacobi<T>(m[0], m[1], m[3], j_right);
  *j_left = rot_mul(rot1, rot_T(*j_right));
}

// A = U diag(S) V^T, row-major NxN, singular values sorted descending.
template <int N, typename T = double>
inline void jacobi_svd(const T* A, T* U, T* S, T* V) {
  const T precision = T(2) * std::numeric_limits<T>::epsilon();
  const T consider_as_zero = std::numeric_limits<T>::min();
  T W[N * N];
  T scale = T(0);
  for (int i = 0; i < N * N; i++) scale = std::max(scale, T(std::fabs(A[i])));
  if (scale == T(0)) scale = T(1);
  for (int i = 0; i < N * N; i++) W[i] = A[i] / scale;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) U[i * N + j] = V[i * N + j] = (i == j) ? T(1) : T(0);
  T max_diag = T(0);
  for (int i = 0; i < N; i++) max_diag = std::max(max_diag, T(std::fabs(W[i * N + i])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 1000) {
    finished = true;
    for (int p = 1; p < N; ++p) {
      for (int q = 0; q < p; ++q) {
        T threshold = std::max(consider_as_zero, precision * max_diag);
        if (std::fabs(W[p * N + q]) > threshold || std::fabs(W[q * N + p]) > threshold) {
          finished = false;
          RotT<T> jl, jr;
          real_2x2_jacobi_svd<T>(W[p * N + p], W[p * N + q], W[q * N + p], W[q * N + q], &jl, &jr);
          apply_left<N, T>(W, p, q, jl);
          apply_right<N, T>(U, p, q, rot_T(jl));
          apply_right<N, T>(W, p, q, jr);
          apply_right<N, T>(V, p, q, jr);
          max_diag = std::max(max_diag, std::max(T(std::fabs(W[p * N + p])), T(std::fabs(W[q * N + q]))));
        }
      }
    }
  }
  for (int i = 0; i < N; i++) {
    T a = std::fabs(W[i * N + i]);
    S[i] = a;
    if (a != T(0)) {
      T f = W[i * N + i] / a;
      for (int r = 0; r < N; r++) U[r * N + i] *= f;
    }
  }
  for (int i = 0; i < N; i++) S[i] *= scale;
  for (int i = 0; i < N; i++) {
    int pos = i;
    T mx = S[i];
    for (int k = i + 1; k < N; k++)
      if (S[k] > mx) {
        mx = S[k];
        pos = k;
      }
    if (mx == T(0)) break;
    if (pos != i) {
      std::swap(S[i], S[pos]);
      for (int r = 0; r < N; r++) {
        std::swap(U[r * N + i], U[r * N + pos]);
        std::swap(V[r * N + i], V[r * N + pos]);
      }
    }
  }
}

// x = V diag(1/S) U^T b over the numerical rank (SVDBase::solve / rank()).
template <int N>
inline void jacobi_svd_solve(const double* A, const double* b, double* x) {
  double U[N * N], S[N], V[N * N];
  jacobi_svd<N, double>(A, U, S, V);
  double thr = std::max(S[0] * (double(N) * DBL_EPSILON), DBL_MIN);
  int rank = N;
  while (rank > 0 && S[rank - 1] < thr) --rank;
  double tmp[N];
  for (int i = 0; i < rank; i++) {
    double acc = 0.0;
    for (int r = 0; r < N; r++) acc += U[r * N + i] * b[r];
    tmp[i] = acc / S[i];
  }
  for (int r = 0; r < N; r++) {
    double acc = 0.0;
    for (int i = 0; i < rank; i++) acc += V[r * N + i] * tmp[i];
    x[r] = acc;
  }
}

// ----------------------------------------------------------------------------------------------
// 3x3 symmetric eigen decomposition, Eigen::SelfAdjointEigenSolver<Matrix3d>::compute restated.
// Reads the lower triangle.  evals ascending; evecs row-major with eigenvectors in columns.
// Returns false on NoConvergence.
inline void make_givens(double p, double q, Rot* r) {
  if (q == 0.0) {
    r->c = p < 0.0 ? -1.0 : 1.0;
    r->s = 0.0;
  } else if (p == 0.0) {
    r->c = 0.0;
    r->s = q < 0.0 ? 1.0 : -1.0;
  } else if (std::fabs(p) > std::fabs(q)) {
    double t = q / p;
    double u = std::sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    r->c = 1.0 / u;
    r->s = -t * r->c;
  } else {
    double t = p / q;
    double u = std::sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    r->s = -1.0 / u;
    r->c = -t * r->s;
  }
}

inline double eigen_hypot(double x, double y) {
  double ax = std::fabs(x), ay = std::fabs(y);
  double p = std::max(ax, ay);
  if (p == 0.0) return 0.0;
  double qp = std::min(ax, ay) / p;
  return p * std::sqrt(1.0 + qp * qp);
}

inline bool self_adjoint_eigen3(const double* A, double* evals, double* evecs) {
  const int n = 3;
  double m[9];
  // lower triangle only
  m[0] = A[0];
  m[3] = A[3];
  m[4] = A[4];
  m[6] = A[6];
  m[7] = A[7];
  m[8] = A[8];
  double scale = 0.0;
  const int lower[6] = {0, 3, 4, 6, 7, 8};
  for (int k = 0; k < 6; k++) scale = std::max(scale, std::fabs(m[lower[k]]));
  if (scale == 0.0) scale = 1.0;
  for (int k = 0; k < 6; k++) m[lower[k]] /= scale;

  double diag[3], sub[2];
  double Q[9];
  // tridiagonalization_inplace_selector<MatrixType,3,false>
  {
    const double tol = DBL_MIN;
    diag[0] = m[0];
    double v1norm2 = m[6] * m[6];
    if (v1norm2 <= tol) {
      diag[1] = m[4];
      diag[2] = m[8];
      sub[0] = m[3];
      sub[1] = m[7];
      for (int i = 0; i < 9; i++) Q[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else {
      double beta = std::sqrt(m[3] * m[3] + v1norm2);
      double inv_beta = 1.0 / beta;
      double m01 = m[3] * inv_beta;
      double m02 = m[6] * inv_beta;
      double q = 2.0 * m01 * m[7] + m02 * (m[8] - m[4]);
      diag[1] = m[4] + m02 * q;
      diag[2] = m[8] - m02 * q;
      sub[0] = beta;
      sub[1] = m[7] - m01 * q;
      Q[0] = 1; Q[1] = 0;   Q[2] = 0;
      Q[3] = 0; Q[4] = m01; Q[5] = m02;
      Q[6] = 0; Q[7] = m02; Q[8] = -m01;
    }
  }
  // computeFromTridiagonal_impl
  const int max_iterations = 30;
  int end = n - 1, start = 0, iter = 0;
  const double consider_as_zero = DBL_MIN;
  const double precision_inv = 1.0 / DBL_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (std::fabs(sub[i]) < consider_as_zero) {
        sub[i] = 0.0;
      } else {
        const double scaled = precision_inv * sub[i];
        if (scaled * scaled <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.0;
      }
    }
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > max_iterations * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    // tridiagonal_qr_step
    {
      double td = (diag[end - 1] - diag[end]) * 0.5;
      double e = sub[end - 1];
      double mu = diag[end];
      if (td == 0.0) {
        mu -= std::fabs(e);
      } else if (e != 0.0) {
        const double e2 = e * e;
        const double h = eigen_hypot(td, e);
        if (e2 == 0.0)
          mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
        else
          mu -= e2 / (td + (td > 0.0 ? h : -h));
      }
      double x = diag[start] - mu;
      double z = sub[start];
      for (int k = start; k < end && z != 0.0; ++k) {
        Rot rot;
        make_givens(x, z, &rot);
        double sdk = rot.s * diag[k] + rot.c * sub[k];
        double dkp1 = rot.s * sub[k] + rot.c * diag[k + 1];
        diag[k] = rot.c * (rot.c * diag[k] - rot.s * sub[k]) - rot.s * (rot.c * sub[k] - rot.s * diag[k + 1]);
        diag[k + 1] = rot.s * sdk + rot.c * dkp1;
        sub[k] = rot.c * sdk - rot.s * dkp1;
        if (k > start) sub[k - 1] = rot.c * sub[k - 1] - rot.s * z;
        x = sub[k];
        if (k < end - 1) {
          z = -rot.s * sub[k + 1];
          sub[k + 1] = rot.c * sub[k + 1];
        }
        apply_right<3, double>(Q, k, k + 1, rot);
      }
    }
  }
  bool ok = iter <= max_iterations * n;
  if (ok) {
    for (int i = 0; i < n - 1; ++i) {
      int k = 0;
      double mn = diag[i];
      for (int j = 1; j < n - i; j++)
        if (diag[i + j] < mn) {
          mn = diag[i + j];
          k = j;
        }
      if (k > 0) {
        std::swap(diag[i], diag[k + i]);
        for (int r = 0; r < 3; r++) std::swap(Q[r * 3 + i], Q[r * 3 + k + i]);
      }
    }
  }
  for (int i = 0; i < 3; i++) evals[i] = diag[i] * scale;
  std::memcpy(evecs, Q, sizeof(Q));
  return ok;
}

// ----------------------------------------------------------------------------------------------
// Matrix3d::inverse(): cofactor expansion along column 0 (Eigen compute_inverse_size3_helper).
inline void inverse3(const double* m, double* r) {
  auto cof = [&](int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
  };
  double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  double det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  double invdet = 1.0 / det;
  r[0] = c00 * invdet;
  r[1] = c10 * invdet;
  r[2] = c20 * invdet;
  r[3] = cof(0, 1) * invdet;
  r[4] = cof(1, 1) * invdet;
  r[5] = cof(2, 1) * invdet;
  r[6] = cof(0, 2) * invdet;
  r[7] = cof(1, 2) * invdet;
  r[8] = cof(2, 2) * invdet;
}

inline void matmul3(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[i * 3 + j] = (a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j]) + a[i * 3 + 2] * b[2 * 3 + j];
}

// ----------------------------------------------------------------------------------------------
// Eigen::LDLT<Matrix<double,6,6>>::solve (lower, symmetric pivoting on the largest |diagonal|).
inline void ldlt_solve6(const double* Ain, const double* b, double* x) {
  const int N = 6;
  double A[36];
  std::memcpy(A, Ain, sizeof(A));
  int transp[N];
  for (int k = 0; k < N; k++) {
    int piv = k;
    double big = std::fabs(A[k * N + k]);
    for (int i = k + 1; i < N; i++)
      if (std::fabs(A[i * N + i]) > big) {
        big = std::fabs(A[i * N + i]);
        piv = i;
      }
    transp[k] = piv;
    if (piv != k) {
      // symmetric swap of rows/cols k and piv on the lower triangle
      int s = N - piv - 1;
      for (int j = 0; j < k; j++) std::swap(A[k * N + j], A[piv * N + j]);
      for (int j = 0; j < s; j++) std::swap(A[(piv + 1 + j) * N + k], A[(piv + 1 + j) * N + piv]);
      std::swap(A[k * N + k], A[piv * N + piv]);
      for (int i = k + 1; i < piv; i++) std::swap(A[i * N + k], A[piv * N + i]);
    }
    int rs = N - k - 1;
    if (k > 0) {
      double temp[N];
      for (int j = 0; j < k; j++) temp[j] = A[j * N + j] * A[k * N + j];
      double acc = 0.0;
      for (int j = 0; j < k; j++) acc += A[k * N + j] * temp[j];
      A[k * N + k] -= acc;
      for (int i = 0; i < rs; i++) {
        double a2 = 0.0;
        for (int j = 0; j < k; j++) a2 += A[(k + 1 + i) * N + j] * temp[j];
        A[(k + 1 + i) * N + k] -= a2;
      }
    }
    double pivot = A[k * N + k];
    if (k == 0 && !(std::fabs(pivot) > 0.0)) {
      for (int j = 0; j < N; j++) transp[j] = j;
      break;
    }
    if (rs > 0 && std::fabs(pivot) > 0.0)
      for (int i = 0; i < rs; i++) A[(k + 1 + i) * N + k] /= pivot;
  }
  // solve: x = P^T L^-T D^-1 L^-1 P b
  double y[N];
  for (int i = 0; i < N; i++) y[i] = b[i];
  for (int k = 0; k < N; k++) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < N; i++) {
    double acc = y[i];
    for (int j = 0; j < i; j++) acc -= A[i * N + j] * y[j];
    y[i] = acc;
  }
  const double tol = DBL_MIN;
  for (int i = 0; i < N; i++) {
    double d = A[i * N + i];
    if (std::fabs(d) > tol)
      y[i] /= d;
    else
      y[i] = 0.0;
  }
  for (int i = N - 1; i >= 0; i--) {
    double acc = y[i];
    for (int j = i + 1; j < N; j++) acc -= A[j * N + i] * y[j];
    y[i] = acc;
  }
  for (int k = N - 1; k >= 0; k--) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < N; i++) x[i] = y[i];
}

// ----------------------------------------------------------------------------------------------
// Matrix3f::eulerAngles(0,1,2) (Eigen/src/Geometry/EulerAngles.h), float arithmetic, row-major R.
inline void euler_angles_012(const float* R, float* res) {
  const float pi = float(M_PI);
  auto c = [&](int r, int cc) { return R[r * 3 + cc]; };
  const int i = 0, j = 1, k = 2;
  res[0] = std::atan2(c(j, k), c(k, k));
  float c2 = std::sqrt(c(i, i) * c(i, i) + c(i, j) * c(i, j));
  if (res[0] > 0.0f) {  // (0,1,2) is an even permutation: fold when the first angle is positive
    res[0] -= pi;
    res[1] = std::atan2(-c(i, k), -c2);
  } else {
    res[1] = std::atan2(-c(i, k), c2);
  }
  float s1 = std::sin(res[0]);
  float c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * c(k, i) - c1 * c(j, i), c1 * c(j, j) - s1 * c(k, j));
  res[0] = -res[0];
  res[1] = -res[1];
  res[2] = -res[2];
}

// Eigen::Transform<float,3,Affine>::rotation() (Eigen/src/Geometry/Transform.h computeRotationScaling):
// polar-decomposition rotation U * V^T from a f32 JacobiSVD of the linear part.  R, out row-major.
inline void affine_rotation_f(const float* R, float* out) {
  float U[9], S[3], V[9];
  jacobi_svd<3, float>(R, U, S, V);
  float UVt[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) UVt[i * 3 + j] = (U[i * 3 + 0] * V[j * 3 + 0] + U[i * 3 + 1] * V[j * 3 + 1]) + U[i * 3 + 2] * V[j * 3 + 2];
  float det = UVt[0] * (UVt[4] * UVt[8] - UVt[5] * UVt[7]) - UVt[1] * (UVt[3] * UVt[8] - UVt[5] * UVt[6]) +
              UVt[2] * (UVt[3] * UVt[7] - UVt[4] * UVt[6]);
  float x = det < 0.0f ? -1.0f : 1.0f;
  for (int r = 0; r < 3; r++) U[r * 3 + 2] *= x;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[i * 3 + j] = (U[i * 3 + 0] * V[j * 3 + 0] + U[i * 3 + 1] * V[j * 3 + 1]) + U[i * 3 + 2] * V[j * 3 + 2];
}

}  // namespace lgs_oracle
