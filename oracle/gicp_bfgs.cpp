// ORACLE (test infrastructure).  pclomp::GeneralizedIterativeClosestPoint restated from
//   GO   = thirdparty/ndt_omp/include/pclomp/gicp_omp_impl.hpp,  GO.h = thirdparty/ndt_omp/include/pclomp/gicp_omp.h
// plus the un-vendored PCL pieces it stands on: pcl::Registration::align (registration.hpp), pcl::transformPointCloud and
// BFGS (pcl/registration/bfgs.h, a port of GSL multimin vector_bfgs2 + linear_minimize: Fletcher's line search with
// bracketing / sectioning and cubic interpolation).  PARITY UNPINNED for those: restated from the published algorithm.
// Transforms are column-major 4x4 floats (Eigen::Matrix4f).  The f / df / fdf sums are EXACT (exactsum.hpp): the
// reference's per-thread partial sums make the low bits depend on the thread count (GO:251,274,291-314), and the
// line search is sensitive to them; an order-independent sum is the reproducible statement of the same quantity.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "exactsum.hpp"
#include "linalg.hpp"
#include "oracle.hpp"

namespace lgs_oracle {

namespace {

void identity_f(float* T) {
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

// Matrix4f * Matrix4f (column-major), coefficient order ((a0*b0 + a1*b1) + a2*b2) + a3*b3
void mul4f(const float* A, const float* B, float* C) {
  float R[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) R[c * 4 + r] = ((A[0 * 4 + r] * B[c * 4 + 0] + A[1 * 4 + r] * B[c * 4 + 1]) + A[2 * 4 + r] * B[c * 4 + 2]) + A[3 * 4 + r] * B[c * 4 + 3];
  std::memcpy(C, R, sizeof(R));
}

// Matrix4f * Vector4f(x, y, z, 1): ((c0*x + c1*y) + c2*z) + c3*1
inline void mul4f_point(const float* T, const P4& p, float out[3]) {
  for (int r = 0; r < 3; r++) out[r] = ((T[0 + r] * p.x + T[4 + r] * p.y) + T[8 + r] * p.z) + T[12 + r] * 1.0f;
}

// ---- BFGS (pcl/registration/bfgs.h) --------------------------------------------------------------------------
struct Functor {
  PclGICP* g;
  double f(const double* x) { return g->functor_f(x); }
  void df(const double* x, double* gr) { g->functor_df(x, gr); }
  void fdf(const double* x, double& f, double* gr) { g->functor_fdf(x, f, gr); }
};

inline double dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; i++) s += a[i] * b[i];
  return s;
}
inline double norm6(const double* a) { return std::sqrt(dot6(a, a)); }

double poly3(const double c[4], double z) { return c[0] + z * (c[1] + z * (c[2] + z * c[3])); }
void check_extremum(const double c[4], double z, double& zmin, double& fmin) {
  const double y = poly3(c, z);
  if (y < fmin) {
    zmin = z;
    fmin = y;
  }
}

// real roots of a x^2 + b x + c (numerically stable form), ascending; returns the count
int solve_quadratic(double a, double b, double c, double* x0, double* x1) {
  if (a == 0) {
    if (b == 0) return 0;
    *x0 = -c / b;
    return 1;
  }
  const double disc = b * b - 4 * a * c;
  if (disc > 0) {
    if (b == 0) {
      const double r = std::sqrt(-c / a);
      *x0 = -r;
      *x1 = r;
    } else {
      const double sgnb = b > 0 ? 1 : -1;
      const double temp = -0.5 * (b + sgnb * std::sqrt(disc));
      const double r1 = temp / a, r2 = c / temp;
      *x0 = std::min(r1, r2);
      *x1 = std::max(r1, r2);
    }
    return 2;
  }
  if (disc == 0) {
    *x0 = *x1 = -0.5 * b / a;
    return 2;
  }
  return 0;
}

double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax, int order) {
  // map [a, b] to [0, 1]
  double y, ymin = (xmin - a) / (b - a), ymax = (xmax - a) / (b - a), fmin;
  if (ymin > ymax) std::swap(ymin, ymax);
  if (order > 2 && !(fpb != fpb) && fpb != std::numeric_limits<double>::infinity()) {
    fpa = fpa * (b - a);
    fpb = fpb * (b - a);
    const double eta = 3 * (fb - fa) - 2 * fpa - fpb;
    const double xi = fpa + fpb - 2 * (fb - fa);
    const double c[4] = {fa, fpa, eta, xi};
    y = ymin;
    fmin = poly3(c, ymin);
    check_extremum(c, ymax, y, fmin);
    double y0, y1;
    const int nroots = solve_quadratic(3 * c[3], 2 * c[2], c[1], &y0, &y1);
    if (nroots == 2) {
      if (y0 > ymin && y0 < ymax) check_extremum(c, y0, y, fmin);
      if (y1 > ymin && y1 < ymax) check_extremum(c, y1, y, fmin);
    } else if (nroots == 1) {
      if (y0 > ymin && y0 < ymax) check_extremum(c, y0, y, fmin);
    }
  } else {
    fpa = fpa * (b - a);
    const double fl = fa + ymin * (fpa + ymin * (fb - fa - fpa));
    const double fh = fa + ymax * (fpa + ymax * (fb - fa - fpa));
    const double c = 2 * (fb - fa - fpa);  // curvature
    y = ymin;
    fmin = fl;
    if (fh < fmin) {
      y = ymax;
      fmin = fh;
    }
    if (c > a) {  // PCL's port tests the curvature against `a` (GSL: c > 0); kept as published
      const double z = -fpa / c;
      if (z > ymin && z < ymax) {
        const double f = fa + z * (fpa + z * (fb - fa - fpa));
        if (f < fmin) {
          y = z;
          fmin = f;
        }
      }
    }
  }
  return a + y * (b - a);
}

struct Bfgs {
  Functor fun;
  BfgsParameters par;
  int iter = 0;
  double f = 0, delta_f = 0, fp0 = 0, pnorm = 0, g0norm = 0;
  double gradient[6], x0[6], dx0[6], dg0[6], g0[6], dx[6], p[6];
  double f_alpha = 0, df_alpha = 0, x_alpha[6], g_alpha[6];
  double f_cache_key = 0, df_cache_key = 0, x_cache_key = 0, g_cache_key = 0;

  double slope() const { return dot6(g_alpha, p); }
  void moveTo(double alpha) {
    if (alpha == x_cache_key) return;
    for (int i = 0; i < 6; i++) x_alpha[i] = x0[i] + alpha * p[i];
    x_cache_key = alpha;
  }
  double applyF(double alpha) {
    if (alpha == f_cache_key) return f_alpha;
    moveTo(alpha);
    f_alpha = fun.f(x_alpha);
    f_cache_key = alpha;
    return f_alpha;
  }
  double applyDF(double alpha) {
    if (alpha == df_cache_key) return df_alpha;
    moveTo(alpha);
    if (alpha != g_cache_key) {
      fun.df(x_alpha, g_alpha);
      g_cache_key = alpha;
    }
    df_alpha = slope();
    df_cache_key = alpha;
    return df_alpha;
  }
  void applyFDF(double alpha, double& fo, double& dfo) {
    if (alpha == f_cache_key && alpha == df_cache_key) {
      fo = f_alpha;
      dfo = df_alpha;
      return;
    }
    if (alpha == f_cache_key || alpha == df_cache_key) {
      fo = applyF(alpha);
      dfo = applyDF(alpha);
      return;
    }
    moveTo(alpha);
    fun.fdf(x_alpha, f_alpha, g_alpha);
    f_cache_key = alpha;
    g_cache_key = alpha;
    df_alpha = slope();
    df_cache_key = alpha;
    fo = f_alpha;
    dfo = df_alpha;
  }
  void updatePosition(double alpha, double* x, double& fo, double* g) {
    double fa, dfa;
    applyFDF(alpha, fa, dfa);
    fo = f_alpha;
    std::memcpy(x, x_alpha, sizeof(x_alpha));
    std::memcpy(g, g_alpha, sizeof(g_alpha));
  }
  void changeDirection() {
    std::memcpy(x_alpha, x0, sizeof(x0));
    x_cache_key = 0;
    f_cache_key = 0;
    f_alpha = f;
    std::memcpy(g_alpha, g0, sizeof(g0));
    g_cache_key = 0;
    df_alpha = slope();
    df_cache_key = 0;
  }

  int minimizeInit(double* x) {
    iter = 0;
    delta_f = 0;
    std::fill(dx, dx + 6, 0.0);
    fun.fdf(x, f, gradient);
    std::memcpy(x0, x, sizeof(x0));
    std::memcpy(g0, gradient, sizeof(g0));
    g0norm = norm6(g0);
    for (int i = 0; i < 6; i++) p[i] = gradient[i] * (-1 / g0norm);
    pnorm = norm6(p);
    fp0 = -g0norm;
    std::memcpy(x_alpha, x0, sizeof(x0));
    x_cache_key = 0;
    f_alpha = f;
    f_cache_key = 0;
    std::memcpy(g_alpha, g0, sizeof(g0));
    g_cache_key = 0;
    df_alpha = slope();
    df_cache_key = 0;
    return BFGS_NOT_STARTED;
  }

  int lineSearch(double rho, double sigma, double tau1, double tau2, double tau3, int order, double alpha1, double& alpha_new) {
    double f0, fp0l, falpha, falpha_prev, fpalpha = 0, fpalpha_prev, delta, alpha_next;
    double alpha = alpha1, alpha_prev = 0.0;
    double a, b, fa, fb, fpa, fpb;
    int i = 0;
    applyFDF(0.0, f0, fp0l);
    falpha_prev = f0;
    fpalpha_prev = fp0l;
    a = 0.0;
    b = alpha;
    fa = f0;
    fb = 0.0;
    fpa = fp0l;
    fpb = 0.0;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    // bracketing
    while (i++ < par.bracket_iters) {
      falpha = applyF(alpha);
      if (falpha > f0 + alpha * rho * fp0l || falpha >= falpha_prev) {  // Fletcher's rho test
        a = alpha_prev;
        fa = falpha_prev;
        fpa = fpalpha_prev;
        b = alpha;
        fb = falpha;
        fpb = nan;
        break;
      }
      fpalpha = applyDF(alpha);
      if (std::fabs(fpalpha) <= -sigma * fp0l) {  // Fletcher's sigma test
        alpha_new = alpha;
        return BFGS_SUCCESS;
      }
      if (fpalpha >= 0) {
        a = alpha;
        fa = falpha;
        fpa = fpalpha;
        b = alpha_prev;
        fb = falpha_prev;
        fpb = fpalpha_prev;
        break;
      }
      delta = alpha - alpha_prev;
      alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, alpha + delta, alpha + tau1 * delta, order);
      alpha_prev = alpha;
      falpha_prev = falpha;
      fpalpha_prev = fpalpha;
      alpha = alpha_next;
    }
    // sectioning of the bracket [a, b]
    while (i++ < par.section_iters) {
      delta = b - a;
      alpha = interpolate(a, fa, fpa, b, fb, fpb, a + tau2 * delta, b - tau3 * delta, order);
      falpha = applyF(alpha);
      if ((a - alpha) * fpa <= std::numeric_limits<double>::epsilon()) return BFGS_NO_PROGRESS;  // roundoff prevents progress
      if (falpha > f0 + rho * alpha * fp0l || falpha >= fa) {
        b = alpha;
        fb = falpha;
        fpb = nan;
      } else {
        fpalpha = applyDF(alpha);
        if (std::fabs(fpalpha) <= -sigma * fp0l) {
          alpha_new = alpha;
          return BFGS_SUCCESS;
        }
        if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
          b = a;
          fb = fa;
          fpb = fpa;
          a = alpha;
          fa = falpha;
          fpa = fpalpha;
        } else {
          a = alpha;
          fa = falpha;
          fpa = fpalpha;
        }
      }
    }
    return BFGS_SUCCESS;
  }

  int minimizeOneStep(double* x) {
    double alpha = 0.0, alpha1;
    const double f0 = f;
    if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0) {
      std::fill(dx, dx + 6, 0.0);
      return BFGS_NO_PROGRESS;
    }
    if (delta_f < 0) {
      const double del = std::max(-delta_f, 10 * std::numeric_limits<double>::epsilon() * std::fabs(f0));
      alpha1 = std::min(1.0, 2.0 * del / (-fp0));
    } else {
      alpha1 = std::fabs(par.step_size);
    }
    const int status = lineSearch(par.rho, par.sigma, par.tau1, par.tau2, par.tau3, par.order, alpha1, alpha);
    if (status != BFGS_SUCCESS) return status;
    updatePosition(alpha, x, f, gradient);
    delta_f = f - f0;
    // BFGS update: p' = g1 - A dx - B dg,  B = dx.g / dx.dg,  A = -(1 + dg.dg / dx.dg) B + dg.g / dx.dg
    for (int i = 0; i < 6; i++) dx0[i] = x[i] - x0[i];
    std::memcpy(dx, dx0, sizeof(dx0));
    for (int i = 0; i < 6; i++) dg0[i] = gradient[i] - g0[i];
    const double dxg = dot6(dx0, gradient), dgg = dot6(dg0, gradient), dxdg = dot6(dx0, dg0), dgnorm = norm6(dg0);
    double A, B;
    if (dxdg != 0) {
      B = dxg / dxdg;
      A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
    } else {
      B = 0;
      A = 0;
    }
    for (int i = 0; i < 6; i++) p[i] = (gradient[i] - A * dx0[i]) - B * dg0[i];
    std::memcpy(g0, gradient, sizeof(g0));
    std::memcpy(x0, x, sizeof(x0));
    g0norm = norm6(g0);
    pnorm = norm6(p);
    const double dir = dot6(p, gradient) > 0 ? -1.0 : 1.0;
    for (int i = 0; i < 6; i++) p[i] *= dir / pnorm;
    pnorm = norm6(p);
    fp0 = dot6(p, g0);
    changeDirection();
    return BFGS_SUCCESS;
  }

  int testGradient(double epsilon) const {
    if (epsilon < 0) return BFGS_NEGATIVE_GRADIENT_EPSILON;
    return norm6(gradient) < epsilon ? BFGS_SUCCESS : BFGS_RUNNING;
  }
};

}  // namespace

PclGICP::PclGICP() {
#ifdef _OPENMP
  num_threads = omp_get_max_threads();
#endif
  identity_f(base_transformation);
  identity_f(transformation);
  identity_f(previous_transformation);
  identity_f(final_transformation);
}

void PclGICP::setInputSource(const P4* p, size_t n) {  // GO.h:139-156: the covariances are dropped
  source.assign(p, p + n);
  source_tree.build(p, n);
  source_covs.clear();
}
void PclGICP::setInputTarget(const P4* p, size_t n) {  // GO.h:164-169
  target.assign(p, p + n);
  target_tree.build(p, n);
  target_covs.clear();
}

// GO:48-122
void PclGICP::computeCovariances(const std::vector<P4>& cloud, const KdTree& tree, std::vector<double>& covs) const {
  const int k = k_correspondences;
  const long n = static_cast<long>(cloud.size());
  if (k > n) return;  // GO:54-58 (PCL_ERROR, covariances left untouched)
  covs.assign(static_cast<size_t>(n) * 9, 0.0);
#pragma omp parallel for num_threads(num_threads) schedule(static)
  for (long i = 0; i < n; i++) {
    std::vector<int32_t> idx(k);
    std::vector<float> d2(k);
    tree.knn(cloud[i], k, idx.data(), d2.data());
    double mean[3] = {0, 0, 0}, c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < k; j++) {
      const P4& pt = cloud[idx[j]];
      mean[0] += pt.x;
      mean[1] += pt.y;
      mean[2] += pt.z;
      c[0] += pt.x * pt.x;  // f32 products widened at the +=
      c[3] += pt.y * pt.x;
      c[4] += pt.y * pt.y;
      c[6] += pt.z * pt.x;
      c[7] += pt.z * pt.y;
      c[8] += pt.z * pt.z;
    }
    for (int a = 0; a < 3; a++) mean[a] /= static_cast<double>(k);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b <= a; b++) {
        c[a * 3 + b] /= static_cast<double>(k);
        c[a * 3 + b] -= mean[a] * mean[b];
        c[b * 3 + a] = c[a * 3 + b];
      }
    double U[9], S[3], V[9];
    jacobi_svd<3, double>(c, U, S, V);
    double* out = &covs[static_cast<size_t>(i) * 9];
    for (int t = 0; t < 9; t++) out[t] = 0.0;
    for (int kk = 0; kk < 3; kk++) {  // cov += v * col * col^T, biggest two singular values -> 1, smallest -> gicp_epsilon
      const double v = kk == 2 ? gicp_epsilon : 1.0;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) out[a * 3 + b] += (v * U[a * 3 + kk]) * U[b * 3 + kk];
    }
  }
}

// GO:518-529.  R = AngleAxisf(x5, Z) * AngleAxisf(x4, Y) * AngleAxisf(x3, X) through f32 quaternions (Eigen).
void PclGICP::applyState(float* t, const double x[6]) const {
  struct Q {
    float w, x, y, z;
  };
  auto axis_q = [](float angle, int axis) {
    const float ha = 0.5f * angle;
    Q q{std::cos(ha), 0.f, 0.f, 0.f};
    const float s = std::sin(ha);
    (axis == 0 ? q.x : axis == 1 ? q.y : q.z) = s;
    return q;
  };
  auto qmul = [](const Q& a, const Q& b) {
    return Q{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
             a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
  };
  const Q q = qmul(qmul(axis_q(static_cast<float>(x[5]), 2), axis_q(static_cast<float>(x[4]), 1)), axis_q(static_cast<float>(x[3]), 0));
  const float tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  const float R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};  // row-major
  float L[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) L[r * 3 + c] = (R[r * 3 + 0] * t[c * 4 + 0] + R[r * 3 + 1] * t[c * 4 + 1]) + R[r * 3 + 2] * t[c * 4 + 2];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[c * 4 + r] = L[r * 3 + c];
  t[12] += static_cast<float>(x[0]);
  t[13] += static_cast<float>(x[1]);
  t[14] += static_cast<float>(x[2]);
  t[15] += 0.0f;
}

// GO:125-178.  R row-major 3x3; g[3..5] = tr(dR^T-style inner products), matricesInnerProd GO.h:314-324
void PclGICP::computeRDerivative(const double x[6], const double R[9], double g[6]) const {
  const double phi = x[3], theta = x[4], psi = x[5];
  const double cphi = std::cos(phi), sphi = std::sin(phi), ctheta = std::cos(theta), stheta = std::sin(theta), cpsi = std::cos(psi), spsi = std::sin(psi);
  double dPhi[9], dTheta[9], dPsi[9];  // row-major
  dPhi[0] = 0; dPhi[3] = 0; dPhi[6] = 0;
  dPhi[1] = sphi * spsi + cphi * cpsi * stheta;
  dPhi[4] = -cpsi * sphi + cphi * spsi * stheta;
  dPhi[7] = cphi * ctheta;
  dPhi[2] = cphi * spsi - cpsi * sphi * stheta;
  dPhi[5] = -cphi * cpsi - sphi * spsi * stheta;
  dPhi[8] = -ctheta * sphi;
  dTheta[0] = -cpsi * stheta;
  dTheta[3] = -spsi * stheta;
  dTheta[6] = -ctheta;
  dTheta[1] = cpsi * ctheta * sphi;
  dTheta[4] = ctheta * sphi * spsi;
  dTheta[7] = -sphi * stheta;
  dTheta[2] = cphi * cpsi * ctheta;
  dTheta[5] = cphi * ctheta * spsi;
  dTheta[8] = -cphi * stheta;
  dPsi[0] = -ctheta * spsi;
  dPsi[3] = cpsi * ctheta;
  dPsi[6] = 0;
  dPsi[1] = -cphi * cpsi - sphi * spsi * stheta;
  dPsi[4] = -cphi * spsi + cpsi * sphi * stheta;
  dPsi[7] = 0;
  dPsi[2] = cpsi * sphi - cphi * spsi * stheta;
  dPsi[5] = sphi * spsi + cphi * cpsi * stheta;
  dPsi[8] = 0;
  auto inner = [&](const double* m1) {  // r += mat1(j, i) * mat2(i, j), i outer, j inner
    double r = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r += m1[j * 3 + i] * R[i * 3 + j];
    return r;
  };
  g[3] = inner(dPhi);
  g[4] = inner(dTheta);
  g[5] = inner(dPsi);
}

// GO:245-275: f32 residual, f32 Mahalanobis product, f32 dot (Eigen 3.4 SSE reduction order (r0 t0 + r2 t2) + (r1 t1 + r3 t3),
// the 4th lane is zero), accumulated in f64, divided by the number of correspondences
double PclGICP::functor_f(const double x[6]) {
  f_calls++;
  float T[16];
  std::memcpy(T, base_transformation, sizeof(T));
  applyState(T, x);
  const size_t m = corr_src.size();
  ExactSum f;
  for (size_t i = 0; i < m; i++) {
    const int si = corr_src[i];
    const P4& ps = output[si];
    const P4& pt = target[corr_tgt[i]];
    float pp[3];
    mul4f_point(T, ps, pp);
    const float res[3] = {pp[0] - pt.x, pp[1] - pt.y, pp[2] - pt.z};
    const float* M = &mahalanobis[static_cast<size_t>(si) * 9];
    float t[3];
    for (int r = 0; r < 3; r++) t[r] = (M[r * 3 + 0] * res[0] + M[r * 3 + 1] * res[1]) + M[r * 3 + 2] * res[2];
    const float ret = (res[0] * t[0] + res[2] * t[2]) + res[1] * t[1];
    f.add(static_cast<double>(ret));
  }
  return f.value() / static_cast<double>(static_cast<int>(m));
}

namespace {
// the per-correspondence term shared by df and fdf: f32 transform and residual, f64 Mahalanobis product
inline void grad_term(const float* T, const P4& ps, const P4& pt, const float* M, double res[3], double temp[3]) {
  float pp[3];
  mul4f_point(T, ps, pp);
  res[0] = pp[0] - pt.x;
  res[1] = pp[1] - pt.y;
  res[2] = pp[2] - pt.z;
  for (int r = 0; r < 3; r++)
    temp[r] = (static_cast<double>(M[r * 3 + 0]) * res[0] + static_cast<double>(M[r * 3 + 1]) * res[1]) + static_cast<double>(M[r * 3 + 2]) * res[2];
}
}  // namespace

// GO:278-330
void PclGICP::functor_df(const double x[6], double g[6]) {
  df_calls++;
  float T[16];
  std::memcpy(T, base_transformation, sizeof(T));
  applyState(T, x);
  const size_t m = corr_src.size();
  ExactSum gs[3], Rs[9];
  for (size_t i = 0; i < m; i++) {
    const int si = corr_src[i];
    const P4& ps = output[si];
    double res[3], temp[3];
    grad_term(T, ps, target[corr_tgt[i]], &mahalanobis[static_cast<size_t>(si) * 9], res, temp);
    float bp[3];
    mul4f_point(base_transformation, ps, bp);  // pp = base_transformation_ * p_src
    for (int a = 0; a < 3; a++) gs[a].add(temp[a]);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) Rs[a * 3 + b].add(static_cast<double>(bp[a]) * temp[b]);
  }
  const double s = 2.0 / static_cast<int>(m);
  double R[9];
  for (int a = 0; a < 6; a++) g[a] = 0;
  for (int a = 0; a < 3; a++) g[a] = gs[a].value() * s;
  for (int t = 0; t < 9; t++) R[t] = Rs[t].value() * s;
  computeRDerivative(x, R, g);
}

// GO:333-367
void PclGICP::functor_fdf(const double x[6], double& f, double g[6]) {
  fdf_calls++;
  float T[16];
  std::memcpy(T, base_transformation, sizeof(T));
  applyState(T, x);
  const size_t m = corr_src.size();
  ExactSum fs, gs[3], Rs[9];
  for (size_t i = 0; i < m; i++) {
    const int si = corr_src[i];
    const P4& ps = output[si];
    double res[3], temp[3];
    grad_term(T, ps, target[corr_tgt[i]], &mahalanobis[static_cast<size_t>(si) * 9], res, temp);
    fs.add((res[0] * temp[0] + res[1] * temp[1]) + res[2] * temp[2]);
    float bp[3];
    mul4f_point(base_transformation, ps, bp);
    for (int a = 0; a < 3; a++) gs[a].add(temp[a]);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) Rs[a * 3 + b].add(static_cast<double>(bp[a]) * temp[b]);
  }
  f = fs.value() / static_cast<double>(static_cast<int>(m));
  const double s = 2.0 / static_cast<int>(m);
  double R[9];
  for (int a = 0; a < 6; a++) g[a] = 0;
  for (int a = 0; a < 3; a++) g[a] = gs[a].value() * s;
  for (int t = 0; t < 9; t++) R[t] = Rs[t].value() * s;
  computeRDerivative(x, R, g);
}

// GO:180-242
bool PclGICP::estimateRigidTransformationBFGS(float* tm) {
  if (corr_src.size() < 4) return false;  // NotEnoughPointsException
  double x[6];
  x[0] = tm[12];
  x[1] = tm[13];
  x[2] = tm[14];
  x[3] = std::atan2(tm[1 * 4 + 2], tm[2 * 4 + 2]);  // atan2(T(2,1), T(2,2)) in f32
  x[4] = std::asin(-tm[0 * 4 + 2]);                 // asin(-T(2,0))
  x[5] = std::atan2(tm[0 * 4 + 1], tm[0 * 4 + 0]);  // atan2(T(1,0), T(0,0))
  const double gradient_tol = 1e-2;
  Bfgs bfgs;
  bfgs.fun.g = this;
  int inner = 0;
  int result = bfgs.minimizeInit(x);
  result = BFGS_RUNNING;
  do {
    inner++;
    result = bfgs.minimizeOneStep(x);
    if (result) break;
    result = bfgs.testGradient(gradient_tol);
  } while (result == BFGS_RUNNING && inner < max_inner_iterations);
  inner_iterations_total += inner;
  if (result == BFGS_NO_PROGRESS || result == BFGS_SUCCESS || inner == max_inner_iterations) {
    identity_f(tm);
    applyState(tm, x);
    return true;
  }
  return false;  // SolverDidntConvergeException
}

// GO:404-474
void PclGICP::update_correspondences(const float* guess) {
  const long N = static_cast<long>(output.size());
  double R[9];  // rotation of transformation_ * guess in f64 (GO:411-417)
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += static_cast<double>(transformation[k * 4 + i]) * static_cast<double>(guess[j * 4 + k]);
      R[i * 3 + j] = s;
    }
  const double dist_threshold = corr_dist_threshold * corr_dist_threshold;
  std::vector<int32_t> match(N, -1);
  mahalanobis.resize(static_cast<size_t>(N) * 9, 0.0f);
#pragma omp parallel for num_threads(num_threads) schedule(static)
  for (long i = 0; i < N; i++) {
    P4 q;
    float qq[3];
    mul4f_point(transformation, output[i], qq);
    q.x = qq[0];
    q.y = qq[1];
    q.z = qq[2];
    q.w = 0;
    int32_t id = -1;
    float d2 = 0;
    if (target_tree.knn(q, 1, &id, &d2) == 0) continue;
    if (static_cast<double>(d2) < dist_threshold) {
      const double* C1 = &source_covs[static_cast<size_t>(i) * 9];
      const double* C2 = &target_covs[static_cast<size_t>(id) * 9];
      double M[9], temp[9], Rt[9], inv[9];
      matmul3(R, C1, M);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) Rt[a * 3 + b] = R[b * 3 + a];
      matmul3(M, Rt, temp);
      for (int t = 0; t < 9; t++) temp[t] += C2[t];
      inverse3(temp, inv);
      for (int t = 0; t < 9; t++) mahalanobis[static_cast<size_t>(i) * 9 + t] = static_cast<float>(inv[t]);
      match[i] = id;
    }
  }
  corr_src.clear();
  corr_tgt.clear();
  for (long i = 0; i < N; i++)
    if (match[i] >= 0) {
      corr_src.push_back(static_cast<int32_t>(i));
      corr_tgt.push_back(match[i]);
    }
}

// pcl::Registration::align (registration.hpp) + GO:370-516
void PclGICP::align(const float* guess, std::vector<P4>* out) {
  f_calls = df_calls = fdf_calls = inner_iterations_total = 0;
  converged = false;
  identity_f(final_transformation);
  identity_f(transformation);
  identity_f(previous_transformation);
  if (target_covs.size() != target.size() * 9) computeCovariances(target, target_tree, target_covs);
  if (source_covs.size() != source.size() * 9) computeCovariances(source, source_tree, source_covs);
  identity_f(base_transformation);
  nr_iterations = 0;
  output.resize(source.size());
  for (size_t i = 0; i < source.size(); i++) output[i] = transform_point(guess, source[i]);  // GO:398
  double delta = 0;
  const bool have_covs = source_covs.size() == source.size() * 9 && target_covs.size() == target.size() * 9;
  while (!converged && have_covs) {
    update_correspondences(guess);
    std::memcpy(previous_transformation, transformation, sizeof(transformation));
    if (!estimateRigidTransformationBFGS(transformation)) break;  // GO:495-499: the exception is caught, the loop ends
    delta = 0.;
    for (int k = 0; k < 4; k++)
      for (int l = 0; l < 4; l++) {
        const double ratio = (k < 3 && l < 3) ? 1. / rotation_epsilon : 1. / transformation_epsilon;
        const double c_delta = ratio * std::abs(previous_transformation[l * 4 + k] - transformation[l * 4 + k]);
        if (c_delta > delta) delta = c_delta;
      }
    nr_iterations++;
    if (nr_iterations >= max_iterations || delta < 1) {
      converged = true;
      std::memcpy(previous_transformation, transformation, sizeof(transformation));
    }
  }
  mul4f(previous_transformation, guess, final_transformation);  // GO:512
  if (out) {
    out->resize(source.size());
    for (size_t i = 0; i < source.size(); i++) (*out)[i] = transform_point(final_transformation, source[i]);
  }
}

double PclGICP::getFitnessScore(double max_range) {
  return fitness_score(target_tree, source.data(), source.size(), final_transformation, max_range, num_threads);
}

}  // namespace lgs_oracle
