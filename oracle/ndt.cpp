// ORACLE (test infrastructure).  pclomp::NormalDistributionsTransform restated from NDT / NDT.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "linalg.hpp"
#include "oracle.hpp"

namespace lgs_oracle {

static void identity16(float* T) {
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

// Eigen::AngleAxis<float>::toRotationMatrix() for a unit axis (Eigen/src/Geometry/AngleAxis.h),
// row-major 3x3 out.  Note the diagonal entry on the axis is (1-c)*1 + c, not a literal 1.
static void angle_axis_unit(float angle, int axis, float* R) {
  float ax[3] = {0, 0, 0};
  ax[axis] = 1.0f;
  float s = std::sin(angle), c = std::cos(angle);
  float sin_axis[3] = {s * ax[0], s * ax[1], s * ax[2]};
  float cos1_axis[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
  float tmp;
  tmp = cos1_axis[0] * ax[1];
  R[0 * 3 + 1] = tmp - sin_axis[2];
  R[1 * 3 + 0] = tmp + sin_axis[2];
  tmp = cos1_axis[0] * ax[2];
  R[0 * 3 + 2] = tmp + sin_axis[1];
  R[2 * 3 + 0] = tmp - sin_axis[1];
  tmp = cos1_axis[1] * ax[2];
  R[1 * 3 + 2] = tmp - sin_axis[0];
  R[2 * 3 + 1] = tmp + sin_axis[0];
  for (int a = 0; a < 3; a++) R[a * 3 + a] = cos1_axis[a] * ax[a] + c;
}

static void matmul3f(const float* a, const float* b, float* c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[i * 3 + j] = (a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j]) + a[i * 3 + 2] * b[2 * 3 + j];
}

// NDT.h:214-231 / NDT:826-829: Translation * AngleAxis(rx,X) * AngleAxis(ry,Y) * AngleAxis(rz,Z), all f32.
void ndt_convert_transform(const double x[6], float* T) {
  float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
  angle_axis_unit(static_cast<float>(x[3]), 0, Rx);
  angle_axis_unit(static_cast<float>(x[4]), 1, Ry);
  angle_axis_unit(static_cast<float>(x[5]), 2, Rz);
  matmul3f(Rx, Ry, Rxy);
  matmul3f(Rxy, Rz, R);
  identity16(T);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[c * 4 + r] = R[r * 3 + c];
  T[12] = static_cast<float>(x[0]);
  T[13] = static_cast<float>(x[1]);
  T[14] = static_cast<float>(x[2]);
}

NDT::NDT() {
  identity16(final_transformation);
#ifdef _OPENMP
  num_threads = omp_get_max_threads();  // NDT:75
#endif
  compute_gauss();
}

void NDT::compute_gauss() {  // NDT:86-93 (and the constructor, NDT:62-69)
  double gauss_c1 = 10 * (1 - outlier_ratio);
  double gauss_c2 = outlier_ratio / std::pow(static_cast<double>(resolution), 3);
  gauss_d3 = -std::log(gauss_c2);
  gauss_d1 = -std::log(gauss_c1 + gauss_c2) - gauss_d3;
  gauss_d2 = -2 * std::log((-std::log(gauss_c1 * std::exp(-0.5) + gauss_c2) - gauss_d3) / gauss_d1);
}

void NDT::init() {  // NDT.h:276-283
  cells.set_leaf_size(resolution, resolution, resolution);
  cells.build(target.data(), target.size());
}

void NDT::setInputTarget(const P4* p, size_t n) {  // NDT.h:122-127
  target.assign(p, p + n);
  tree_dirty = true;
  init();
}

void NDT::setInputSource(const P4* p, size_t n) {
  source.assign(p, p + n);
  have_source = true;
}

void NDT::setResolution(float r) {  // NDT.h:132-142
  if (resolution != r) {
    resolution = r;
    if (have_source) init();
  }
}

// NDT:288-394
void NDT::computeAngleDerivatives(const double p[6]) {
  double cx, cy, cz, sx, sy, sz;
  if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
  if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
  if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }

  const double J[8][3] = {
      {(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},  // a
      {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},    // b
      {(-sy * cz), sy * sz, cy},                                           // c
      {sx * cy * cz, (-sx * cy * sz), sx * sy},                            // d
      {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},                         // e
      {(-cy * sz), (-cy * cz), 0},                                         // f
      {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},            // g
      {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};            // h
  const double Hh[15][3] = {
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},     // a2
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)},  // a3
      {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},                        // b2
      {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},                        // b3
      {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},            // c2
      {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},            // c3
      {(-cy * cz), (cy * sz), (sy)},                                       // d1
      {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},                        // d2
      {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},                       // d3
      {(sy * sz), (sy * cz), 0},                                           // e1
      {(-sx * cy * sz), (-sx * cy * cz), 0},                               // e2
      {(cx * cy * sz), (cx * cy * cz), 0},                                 // e3
      {(-cy * cz), (cy * sz), 0},                                          // f1
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},           // f2
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};          // f3
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 3; c++) {
      j_ang_d[r][c] = J[r][c];
      j_ang[r][c] = static_cast<float>(J[r][c]);
    }
  for (int r = 0; r < 8; r++) j_ang[r][3] = 0.0f;
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 3; c++) {
      h_ang_d[r][c] = Hh[r][c];
      h_ang[r][c] = static_cast<float>(Hh[r][c]);
    }
  for (int r = 0; r < 16; r++) h_ang[r][3] = 0.0f;
  for (int c = 0; c < 3; c++) h_ang[15][c] = 0.0f;
}

namespace {

// f32 point derivative tables of one source point (NDT:397-439): the 8 non-trivial entries of
// point_gradient_ (4x6) and the 6 vectors a..f of point_hessian_ (24x6).
struct PointDerivF {
  float J[4][6];      // point_gradient_
  float Hv[6][4];     // a,b,c,d,e,f as 4-vectors (4th = 0)
};

inline void point_derivatives_f(const float j_ang[8][4], const float h_ang[16][4], const double x[3], PointDerivF* o) {
  const float x4[4] = {static_cast<float>(x[0]), static_cast<float>(x[1]), static_cast<float>(x[2]), 0.0f};
  float xj[8], xh[16];
  // j_ang * x4 (column-major GEMV: accumulate column by column)
  for (int r = 0; r < 8; r++) xj[r] = ((j_ang[r][0] * x4[0] + j_ang[r][1] * x4[1]) + j_ang[r][2] * x4[2]) + j_ang[r][3] * x4[3];
  for (int r = 0; r < 16; r++) xh[r] = ((h_ang[r][0] * x4[0] + h_ang[r][1] * x4[1]) + h_ang[r][2] * x4[2]) + h_ang[r][3] * x4[3];
  std::memset(o->J, 0, sizeof(o->J));
  o->J[0][0] = o->J[1][1] = o->J[2][2] = 1.0f;
  o->J[1][3] = xj[0];
  o->J[2][3] = xj[1];
  o->J[0][4] = xj[2];
  o->J[1][4] = xj[3];
  o->J[2][4] = xj[4];
  o->J[0][5] = xj[5];
  o->J[1][5] = xj[6];
  o->J[2][5] = xj[7];
  const float a[4] = {0, xh[0], xh[1], 0}, b[4] = {0, xh[2], xh[3], 0}, c[4] = {0, xh[4], xh[5], 0};
  const float d[4] = {xh[6], xh[7], xh[8], 0}, e[4] = {xh[9], xh[10], xh[11], 0}, f[4] = {xh[12], xh[13], xh[14], 0};
  std::memcpy(o->Hv[0], a, 16);
  std::memcpy(o->Hv[1], b, 16);
  std::memcpy(o->Hv[2], c, 16);
  std::memcpy(o->Hv[3], d, 16);
  std::memcpy(o->Hv[4], e, 16);
  std::memcpy(o->Hv[5], f, 16);
}

// which of a..f sits in block (i,j), i,j in 3..5 (NDT:429-437)
const int kHessBlock[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};

}  // namespace

// NDT:483-536.  All products f32, accumulation into f64 outputs.
static double update_derivatives_f(double g[6], double H[36], const PointDerivF& pd, const double x_trans[3], const double c_inv[9],
                                   bool compute_hessian, double gauss_d1, double gauss_d2_d) {
  const float x4[4] = {static_cast<float>(x_trans[0]), static_cast<float>(x_trans[1]), static_cast<float>(x_trans[2]), 0.0f};
  float C[4][4];
  std::memset(C, 0, sizeof(C));
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) C[r][c] = static_cast<float>(c_inv[r * 3 + c]);
  const float gauss_d2 = static_cast<float>(gauss_d2_d);

  // x_trans4 * c_inv4  (1x4)
  float xC[4];
  for (int c = 0; c < 4; c++) xC[c] = ((x4[0] * C[0][c] + x4[1] * C[1][c]) + x4[2] * C[2][c]) + x4[3] * C[3][c];
  const float q = ((x4[0] * xC[0] + x4[1] * xC[1]) + x4[2] * xC[2]) + x4[3] * xC[3];
  // exp() of an f32 argument; evaluated in f64 and rounded (== correctly rounded expf up to double rounding)
  float e_x_cov_x = static_cast<float>(std::exp(static_cast<double>(-gauss_d2 * q * 0.5f)));
  const float score_inc = static_cast<float>(-gauss_d1 * e_x_cov_x);
  e_x_cov_x = gauss_d2 * e_x_cov_x;
  if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) return 0;
  e_x_cov_x = static_cast<float>(e_x_cov_x * gauss_d1);

  // c_inv4 * point_gradient4 (4x6)
  float CJ[4][6];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 6; c++) CJ[r][c] = ((C[r][0] * pd.J[0][c] + C[r][1] * pd.J[1][c]) + C[r][2] * pd.J[2][c]) + C[r][3] * pd.J[3][c];
  float xCJ[6];
  for (int c = 0; c < 6; c++) xCJ[c] = ((x4[0] * CJ[0][c] + x4[1] * CJ[1][c]) + x4[2] * CJ[2][c]) + x4[3] * CJ[3][c];
  for (int c = 0; c < 6; c++) g[c] += static_cast<double>(e_x_cov_x * xCJ[c]);

  if (compute_hessian) {
    float JCJ[6][6];  // point_gradient4^T * (c_inv4 * point_gradient4)
    for (int a = 0; a < 6; a++)
      for (int b = 0; b < 6; b++) JCJ[a][b] = ((pd.J[0][a] * CJ[0][b] + pd.J[1][a] * CJ[1][b]) + pd.J[2][a] * CJ[2][b]) + pd.J[3][a] * CJ[3][b];
    for (int i = 0; i < 6; i++) {
      float xCH[6] = {0, 0, 0, 0, 0, 0};  // x_trans4_x_c_inv4 * point_hessian_.block<4,6>(i*4,0)
      if (i >= 3)
        for (int j = 3; j < 6; j++) {
          const float* v = pd.Hv[kHessBlock[i - 3][j - 3]];
          xCH[j] = ((xC[0] * v[0] + xC[1] * v[1]) + xC[2] * v[2]) + xC[3] * v[3];
        }
      for (int j = 0; j < 6; j++) {
        const float term = e_x_cov_x * (-gauss_d2 * xCJ[i] * xCJ[j] + xCH[j] + JCJ[j][i]);
        H[i * 6 + j] += static_cast<double>(term);
      }
    }
  }
  return score_inc;
}

// NDT:179-285
double NDT::computeDerivatives(double g[6], double H[36], const std::vector<P4>& trans_cloud, const double p[6], bool compute_hessian) {
  stats.derivative_evals++;
  const size_t n = source.size();
  std::fill(g, g + 6, 0.0);
  std::fill(H, H + 36, 0.0);
  std::vector<double> scores(n, 0.0), grads(n * 6, 0.0), hess(n * 36, 0.0);

  computeAngleDerivatives(p);

#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (long idx = 0; idx < static_cast<long>(n); idx++) {
    const P4& x_trans_pt = trans_cloud[idx];
    const Leaf* nb[27];
    int cnt = cells.neighborhood(x_trans_pt, search_method, nb);
    double score_pt = 0;
    double g_pt[6] = {0, 0, 0, 0, 0, 0};
    double H_pt[36];
    std::fill(H_pt, H_pt + 36, 0.0);
    const double x[3] = {source[idx].x, source[idx].y, source[idx].z};
    PointDerivF pd;
    if (cnt) point_derivatives_f(j_ang, h_ang, x, &pd);
    for (int k = 0; k < cnt; k++) {
      const Leaf* cell = nb[k];
      double x_trans[3] = {x_trans_pt.x - cell->mean[0], x_trans_pt.y - cell->mean[1], x_trans_pt.z - cell->mean[2]};
      score_pt += update_derivatives_f(g_pt, H_pt, pd, x_trans, cell->icov, compute_hessian, gauss_d1, gauss_d2);
    }
    scores[idx] = score_pt;
    std::copy(g_pt, g_pt + 6, &grads[idx * 6]);
    std::copy(H_pt, H_pt + 36, &hess[idx * 36]);
  }
  // NDT:277-282: serial, index order
  double score = 0;
  for (size_t i = 0; i < n; i++) {
    score += scores[i];
    for (int k = 0; k < 6; k++) g[k] += grads[i * 6 + k];
    for (int k = 0; k < 36; k++) H[k] += hess[i * 36 + k];
  }
  static const bool eval_trace = getenv("LGS_NDT_EVAL_TRACE") != nullptr;  // tests/diag_ndt_eval_diff.py: pose, transform, sums
  if (eval_trace) {
    fprintf(stderr, "EV mode %d P", compute_hessian ? 0 : 1);
    for (int i = 0; i < 6; i++) fprintf(stderr, " %a", p[i]);
    fprintf(stderr, " T");
    for (int i = 0; i < 12; i++) fprintf(stderr, " %a", static_cast<double>(final_transformation[i]));
    fprintf(stderr, " | S %a", score);
    for (int k = 0; k < 6; k++) fprintf(stderr, " %a", g[k]);
    if (compute_hessian) {
      for (int r = 0; r < 6; r++)
        for (int c = r; c < 6; c++) fprintf(stderr, " %a", H[r * 6 + c]);
      for (int r = 1; r < 6; r++)
        for (int c = 0; c < r; c++) fprintf(stderr, " %a", H[r * 6 + c]);
    }
    fprintf(stderr, "\n");
  }
  return score;
}

// NDT:539-644 (computeHessian + updateHessian, f64 formulas NDT:443-480 for the point derivatives)
void NDT::computeHessian(double H[36], const std::vector<P4>& trans_cloud) {
  stats.hessian_recomputes++;
  std::fill(H, H + 36, 0.0);
  const size_t n = source.size();
  for (size_t idx = 0; idx < n; idx++) {
    const P4& x_trans_pt = trans_cloud[idx];
    const Leaf* nb[27];
    int cnt = cells.neighborhood(x_trans_pt, search_method, nb);
    if (!cnt) continue;
    const double x[3] = {source[idx].x, source[idx].y, source[idx].z};
    double J[3][6];
    std::memset(J, 0, sizeof(J));
    J[0][0] = J[1][1] = J[2][2] = 1.0;
    auto dotj = [&](int r) { return sum3d(x[0] * j_ang_d[r][0], x[1] * j_ang_d[r][1], x[2] * j_ang_d[r][2]); };
    auto doth = [&](int r) { return sum3d(x[0] * h_ang_d[r][0], x[1] * h_ang_d[r][1], x[2] * h_ang_d[r][2]); };
    J[1][3] = dotj(0); J[2][3] = dotj(1);
    J[0][4] = dotj(2); J[1][4] = dotj(3); J[2][4] = dotj(4);
    J[0][5] = dotj(5); J[1][5] = dotj(6); J[2][5] = dotj(7);
    const double vec[6][3] = {{0, doth(0), doth(1)},       {0, doth(2), doth(3)},          {0, doth(4), doth(5)},
                              {doth(6), doth(7), doth(8)}, {doth(9), doth(10), doth(11)}, {doth(12), doth(13), doth(14)}};
    for (int k = 0; k < cnt; k++) {
      const Leaf* cell = nb[k];
      const double xt[3] = {x_trans_pt.x - cell->mean[0], x_trans_pt.y - cell->mean[1], x_trans_pt.z - cell->mean[2]};
      const double* C = cell->icov;
      auto Cv = [&](const double* v, double* o) {
        for (int r = 0; r < 3; r++) o[r] = sum3d(C[r * 3 + 0] * v[0], C[r * 3 + 1] * v[1], C[r * 3 + 2] * v[2]);
      };
      auto dot3 = [&](const double* a, const double* b) { return sum3d(a[0] * b[0], a[1] * b[1], a[2] * b[2]); };
      double Cx[3];
      Cv(xt, Cx);
      double e_x_cov_x = gauss_d2 * std::exp(-gauss_d2 * dot3(xt, Cx) / 2);
      if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) continue;
      e_x_cov_x *= gauss_d1;
      double CJ[6][3], xCJ[6];
      for (int i = 0; i < 6; i++) {
        const double col[3] = {J[0][i], J[1][i], J[2][i]};
        Cv(col, CJ[i]);
        xCJ[i] = dot3(xt, CJ[i]);
      }
      for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 6; j++) {
          double xCH = 0.0;
          {
            double hv[3] = {0, 0, 0};
            if (i >= 3 && j >= 3) {
              const double* v = vec[kHessBlock[i - 3][j - 3]];
              hv[0] = v[0]; hv[1] = v[1]; hv[2] = v[2];
            }
            double Ch[3];
            Cv(hv, Ch);
            xCH = dot3(xt, Ch);
          }
          const double colj[3] = {J[0][j], J[1][j], J[2][j]};
          H[i * 6 + j] += e_x_cov_x * (-gauss_d2 * xCJ[i] * xCJ[j] + xCH + dot3(colj, CJ[i]));
        }
      }
    }
  }
}

// NDT:647-685
static bool update_interval_mt(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t,
                               double g_t) {
  if (f_t > f_l) {
    a_u = a_t; f_u = f_t; g_u = g_t;
    return false;
  } else if (g_t * (a_l - a_t) > 0) {
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  } else if (g_t * (a_l - a_t) < 0) {
    a_u = a_l; f_u = f_l; g_u = g_l;
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  }
  return true;
}

// NDT:688-768
static double trial_value_selection_mt(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t,
                                       double g_t) {
  if (f_t > f_l) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (std::fabs(g_t) <= std::fabs(g_l)) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
    return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
  }
  double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
  double w = std::sqrt(z * z - g_t * g_u);
  return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

// NDT.h:430-447
static inline double psi_mt(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
static inline double dpsi_mt(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

static inline double dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; i++) s += a[i] * b[i];
  return s;
}

// NDT:771-931
double NDT::computeStepLengthMT(const double x[6], double step_dir[6], double step_init, double step_max, double step_min, double& score,
                                double g[6], double H[36], std::vector<P4>& trans_cloud) {
  double phi_0 = -score;
  double d_phi_0 = -dot6(g, step_dir);
  double x_t[6];
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0;
    d_phi_0 *= -1;
    for (int i = 0; i < 6; i++) step_dir[i] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double a_l = 0, a_u = 0;
  double f_l = psi_mt(a_l, phi_0, phi_0, d_phi_0, mu);
  double g_l = dpsi_mt(d_phi_0, d_phi_0, mu);
  double f_u = psi_mt(a_u, phi_0, phi_0, d_phi_0, mu);
  double g_u = dpsi_mt(d_phi_0, d_phi_0, mu);
  bool interval_converged = (step_max - step_min) < 0, open_interval = true;
  double a_t = step_init;
  a_t = std::min(a_t, step_max);
  a_t = std::max(a_t, step_min);
  for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
  ndt_convert_transform(x_t, final_transformation);
  for (size_t i = 0; i < source.size(); i++) trans_cloud[i] = transform_point(final_transformation, source[i]);
  score = computeDerivatives(g, H, trans_cloud, x_t, true);
  double phi_t = -score;
  double d_phi_t = -dot6(g, step_dir);
  double psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu);
  double d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);

  while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    stats.line_search_trials++;
    if (open_interval)
      a_t = trial_value_selection_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else
      a_t = trial_value_selection_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
    ndt_convert_transform(x_t, final_transformation);
    for (size_t i = 0; i < source.size(); i++) trans_cloud[i] = transform_point(final_transformation, source[i]);
    score = computeDerivatives(g, H, trans_cloud, x_t, false);
    phi_t = -score;
    d_phi_t = -dot6(g, step_dir);
    psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu);
    d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = false;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
      g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
      g_u = g_u + mu * d_phi_0;
    }
    if (open_interval)
      interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else
      interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    step_iterations++;
  }
  if (step_iterations) computeHessian(H, trans_cloud);
  return a_t;
}

// pcl::Registration::align shell (PCL registration.hpp) + NDT:80-171
void NDT::align(const float* guess, std::vector<P4>* output_opt) {
  stats = NdtStats();
  std::vector<P4> local;
  std::vector<P4>& output = output_opt ? *output_opt : local;
  output = source;  // align(): copy the input, data[3] = 1 (our w carries intensity; the 4th homogeneous lane is implicit)
  identity16(final_transformation);
  converged = false;
  nr_iterations = 0;
  compute_gauss();

  bool guess_is_identity = true;
  for (int i = 0; i < 16; i++)
    if (guess[i] != ((i % 5 == 0) ? 1.0f : 0.0f)) guess_is_identity = false;
  if (!guess_is_identity) {  // NDT:95-101
    std::memcpy(final_transformation, guess, sizeof(float) * 16);
    for (size_t i = 0; i < output.size(); i++) output[i] = transform_point(guess, output[i]);
  }
  // NDT:103-111
  float R[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[r * 3 + c] = final_transformation[c * 4 + r];
  float Rpolar[9], rot[3];
  affine_rotation_f(R, Rpolar);  // Transform<float,3,Affine>::rotation() at NDT:109 is the polar rotation, not linear()
  euler_angles_012(Rpolar, rot);
  double p[6] = {final_transformation[12], final_transformation[13], final_transformation[14], rot[0], rot[1], rot[2]};
  double delta_p[6], g[6], H[36];
  double score = computeDerivatives(g, H, output, p, true);  // NDT:119
  const double n_in = static_cast<double>(source.size());

  while (!converged) {
    double neg_g[6];
    for (int i = 0; i < 6; i++) neg_g[i] = -g[i];
    jacobi_svd_solve<6>(H, neg_g, delta_p);  // NDT:127-129
    double delta_p_norm = std::sqrt(dot6(delta_p, delta_p));
    if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {  // NDT:134-139
      trans_probability = score / n_in;
      converged = delta_p_norm == delta_p_norm;
      return;
    }
    for (int i = 0; i < 6; i++) delta_p[i] /= delta_p_norm;  // normalize()
    delta_p_norm = computeStepLengthMT(p, delta_p, delta_p_norm, step_size, transformation_epsilon / 2, score, g, H, output);
    for (int i = 0; i < 6; i++) delta_p[i] *= delta_p_norm;
    for (int i = 0; i < 6; i++) p[i] = p[i] + delta_p[i];
    if (nr_iterations > max_iterations || (nr_iterations && (std::fabs(delta_p_norm) < transformation_epsilon))) converged = true;
    nr_iterations++;
  }
  trans_probability = score / n_in;
}

double NDT::getFitnessScore(double max_range) {
  if (tree_dirty) {
    target_tree.build(target.data(), target.size());
    tree_dirty = false;
  }
  return fitness_score(target_tree, source.data(), source.size(), final_transformation, max_range, num_threads);
}

// NDT:934-982
double NDT::calculateScore(const std::vector<P4>& trans_cloud) const {
  double score = 0;
  for (size_t idx = 0; idx < trans_cloud.size(); idx++) {
    const P4& x_trans_pt = trans_cloud[idx];
    const Leaf* nb[27];
    int cnt = cells.neighborhood(x_trans_pt, search_method, nb);
    for (int k = 0; k < cnt; k++) {
      const Leaf* cell = nb[k];
      const double xt[3] = {x_trans_pt.x - cell->mean[0], x_trans_pt.y - cell->mean[1], x_trans_pt.z - cell->mean[2]};
      const double* C = cell->icov;
      double Cx[3];
      for (int r = 0; r < 3; r++) Cx[r] = sum3d(C[r * 3 + 0] * xt[0], C[r * 3 + 1] * xt[1], C[r * 3 + 2] * xt[2]);
      double e_x_cov_x = std::exp(-gauss_d2 * sum3d(xt[0] * Cx[0], xt[1] * Cx[1], xt[2] * Cx[2]) / 2);
      double score_inc = -gauss_d1 * e_x_cov_x - gauss_d3;
      score += score_inc / cnt;
    }
  }
  return score / static_cast<double>(trans_cloud.size());
}

}  // namespace lgs_oracle
