// ORACLE (test infrastructure).  fast_gicp::FastGICP + fast_gicp::LsqRegistration restated from FG / LSQ.
// Transforms are row-major 4x4 doubles (Eigen::Isometry3d in the reference).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "linalg.hpp"
#include "oracle.hpp"

namespace lgs_oracle {

static void identity_d(double* T) {
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

// Isometry3d * Isometry3d (Eigen transform_transform_product_impl): linear = L*L', translation = L*t' + t.
static void iso_mul(const double* A, const double* B, double* C) {
  double R[16];
  identity_d(R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 4 + j] = (A[i * 4 + 0] * B[0 * 4 + j] + A[i * 4 + 1] * B[1 * 4 + j]) + A[i * 4 + 2] * B[2 * 4 + j];
  for (int i = 0; i < 3; i++) R[i * 4 + 3] = ((A[i * 4 + 0] * B[0 * 4 + 3] + A[i * 4 + 1] * B[1 * 4 + 3]) + A[i * 4 + 2] * B[2 * 4 + 3]) + A[i * 4 + 3];
  std::memcpy(C, R, sizeof(R));
}

// so3/so3.hpp:58-77 (so3_exp) followed by Eigen::Quaterniond::toRotationMatrix
static void so3_exp_matrix(const double* omega, double* R /*row-major 3x3*/) {
  double theta_sq = (omega[0] * omega[0] + omega[1] * omega[1]) + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double theta = std::sqrt(theta_sq);
    double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  const double w = real_factor, x = imag_factor * omega[0], y = imag_factor * omega[1], z = imag_factor * omega[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

FastGICP::FastGICP() {
#ifdef _OPENMP
  num_threads = omp_get_max_threads();  // FG:10-14
#endif
  std::memset(final_hessian, 0, sizeof(final_hessian));
  for (int i = 0; i < 6; i++) final_hessian[i * 7] = 1.0;  // LSQ:21
  for (int i = 0; i < 16; i++) final_transformation[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

// The reference caches on shared_ptr identity (FG:72-90); every call here carries a fresh buffer,
// i.e. the "new pointer" branch: reset cloud, kd-tree input and covariances.
void FastGICP::setInputSource(const P4* p, size_t n) {
  source = std::make_shared<Cloud>();
  source->pts.assign(p, p + n);
  source->tree.build(p, n);
}
void FastGICP::setInputTarget(const P4* p, size_t n) {
  target = std::make_shared<Cloud>();
  target->pts.assign(p, p + n);
  target->tree.build(p, n);
}
void FastGICP::swapSourceAndTarget() {  // FG:50-57
  source.swap(target);
  correspondences.clear();
  sq_distances.clear();
}
void FastGICP::clearSource() { source.reset(); }
void FastGICP::clearTarget() { target.reset(); }

// FG:241-298
void FastGICP::calculate_covariances(Cloud& c) {
  const int k = k_correspondences;
  const long n = static_cast<long>(c.pts.size());
  c.covs.assign(static_cast<size_t>(n) * 9, 0.0);
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (long i = 0; i < n; i++) {
    std::vector<int32_t> idx(k);
    std::vector<float> d2(k);
    int found = c.tree.knn(c.pts[i], k, idx.data(), d2.data());
    // neighbors is 4 x k; columns beyond `found` are uninitialised in the reference, zero here
    std::vector<double> nb(static_cast<size_t>(k) * 3, 0.0);
    for (int j = 0; j < found; j++) {
      nb[j * 3 + 0] = c.pts[idx[j]].x;
      nb[j * 3 + 1] = c.pts[idx[j]].y;
      nb[j * 3 + 2] = c.pts[idx[j]].z;
    }
    double mean[3] = {0, 0, 0};
    for (int a = 0; a < 3; a++) {
      double s = 0;
      for (int j = 0; j < k; j++) s += nb[j * 3 + a];
      mean[a] = s / k;
    }
    for (int j = 0; j < k; j++)
      for (int a = 0; a < 3; a++) nb[j * 3 + a] -= mean[a];
    double cov[9];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double s = 0;
        for (int j = 0; j < k; j++) s += nb[j * 3 + a] * nb[j * 3 + b];
        cov[a * 3 + b] = s / k;
      }
    double* out = &c.covs[static_cast<size_t>(i) * 9];
    if (regularization == REG_NONE) {
      std::copy(cov, cov + 9, out);
    } else if (regularization == REG_FROBENIUS) {
      const double lambda = 1e-3;
      double C[9], Cinv[9];
      for (int t = 0; t < 9; t++) C[t] = cov[t] + ((t % 4 == 0) ? lambda : 0.0);
      inverse3(C, Cinv);
      double nrm = 0;
      for (int t = 0; t < 9; t++) nrm += Cinv[t] * Cinv[t];
      nrm = std::sqrt(nrm);
      for (int t = 0; t < 9; t++) Cinv[t] /= nrm;
      inverse3(Cinv, out);
    } else {
      double U[9], S[3], V[9], values[3];
      jacobi_svd<3, double>(cov, U, S, V);
      if (regularization == REG_PLANE) {
        values[0] = 1; values[1] = 1; values[2] = 1e-3;
      } else if (regularization == REG_MIN_EIG) {
        for (int a = 0; a < 3; a++) values[a] = std::max(S[a], 1e-3);
      } else {  // NORMALIZED_MIN_EIG
        double mx = std::max(S[0], std::max(S[1], S[2]));
        for (int a = 0; a < 3; a++) values[a] = std::max(S[a] / mx, 1e-3);
      }
      double UD[9];
      for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) UD[r * 3 + cc] = U[r * 3 + cc] * values[cc];
      for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) out[r * 3 + cc] = (UD[r * 3 + 0] * V[cc * 3 + 0] + UD[r * 3 + 1] * V[cc * 3 + 1]) + UD[r * 3 + 2] * V[cc * 3 + 2];
    }
  }
}

// FG:115-152
void FastGICP::update_correspondences(const double T[16]) {
  const long n = static_cast<long>(source->pts.size());
  float Tf[16];
  for (int i = 0; i < 16; i++) Tf[i] = static_cast<float>(T[i]);
  correspondences.resize(n);
  sq_distances.resize(n);
  mahalanobis.resize(static_cast<size_t>(n) * 9);
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (long i = 0; i < n; i++) {
    const P4& s = source->pts[i];
    P4 pt;  // trans_f * input (Isometry3f * Vector4f: column-by-column GEMV)
    pt.x = ((Tf[0] * s.x + Tf[1] * s.y) + Tf[2] * s.z) + Tf[3] * 1.0f;
    pt.y = ((Tf[4] * s.x + Tf[5] * s.y) + Tf[6] * s.z) + Tf[7] * 1.0f;
    pt.z = ((Tf[8] * s.x + Tf[9] * s.y) + Tf[10] * s.z) + Tf[11] * 1.0f;
    pt.w = 0;
    int32_t id = -1;
    float d2 = 0;
    target->tree.knn(pt, 1, &id, &d2);
    sq_distances[i] = d2;
    correspondences[i] = d2 < corr_dist_threshold * corr_dist_threshold ? id : -1;
    if (correspondences[i] < 0) continue;
    const double* cov_A = &source->covs[static_cast<size_t>(i) * 9];
    const double* cov_B = &target->covs[static_cast<size_t>(id) * 9];
    // RCR = cov_B + T cov_A T^T; the 4th row/column of the reference's Matrix4d stay zero, (3,3) is
    // patched to 1 before inverse() and back to 0 after it, so only the 3x3 block carries data.
    double RC[9], RCR[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) RC[r * 3 + c] = (T[r * 4 + 0] * cov_A[0 * 3 + c] + T[r * 4 + 1] * cov_A[1 * 3 + c]) + T[r * 4 + 2] * cov_A[2 * 3 + c];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        RCR[r * 3 + c] = cov_B[r * 3 + c] + ((RC[r * 3 + 0] * T[c * 4 + 0] + RC[r * 3 + 1] * T[c * 4 + 1]) + RC[r * 3 + 2] * T[c * 4 + 2]);
    inverse3(RCR, &mahalanobis[static_cast<size_t>(i) * 9]);
  }
}

namespace {
struct Accum {
  double H[36];
  double b[6];
  double e;
  char pad[64];
};
}  // namespace

// FG:155-211 / FG:214-237.  want_Hb=false is compute_error.
static double gicp_accumulate(const FastGICP& g, const double T[16], double* H, double* b, bool want_Hb) {
  const long n = static_cast<long>(g.source->pts.size());
  const int nt = std::max(1, g.num_threads);
  std::vector<Accum> acc(nt);
  for (auto& a : acc) std::memset(&a, 0, sizeof(Accum));
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    Accum& A = acc[tid];
#pragma omp for schedule(static)
    for (long i = 0; i < n; i++) {
      int ti = g.correspondences[i];
      if (ti < 0) continue;
      const P4& a = g.source->pts[i];
      const P4& bb = g.target->pts[ti];
      const double mA[3] = {a.x, a.y, a.z};
      const double mB[3] = {bb.x, bb.y, bb.z};
      double tA[3], err[3];
      for (int r = 0; r < 3; r++) tA[r] = ((T[r * 4 + 0] * mA[0] + T[r * 4 + 1] * mA[1]) + T[r * 4 + 2] * mA[2]) + T[r * 4 + 3] * 1.0;
      for (int r = 0; r < 3; r++) err[r] = mB[r] - tA[r];
      const double* M = &g.mahalanobis[static_cast<size_t>(i) * 9];
      double eM[3];
      for (int c = 0; c < 3; c++) eM[c] = (err[0] * M[0 * 3 + c] + err[1] * M[1 * 3 + c]) + err[2] * M[2 * 3 + c];
      A.e += (eM[0] * err[0] + eM[1] * err[1]) + eM[2] * err[2];
      if (!want_Hb) continue;
      // dtdx0 = [skewd(tA), -I]  (3x6 live rows of the reference's 4x6)
      double J[3][6] = {{0, -tA[2], tA[1], -1, 0, 0}, {tA[2], 0, -tA[0], 0, -1, 0}, {-tA[1], tA[0], 0, 0, 0, -1}};
      double JtM[6][3];
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 3; c++) JtM[r][c] = (J[0][r] * M[0 * 3 + c] + J[1][r] * M[1 * 3 + c]) + J[2][r] * M[2 * 3 + c];
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) A.H[r * 6 + c] += (JtM[r][0] * J[0][c] + JtM[r][1] * J[1][c]) + JtM[r][2] * J[2][c];
        A.b[r] += (JtM[r][0] * err[0] + JtM[r][1] * err[1]) + JtM[r][2] * err[2];
      }
    }
  }
  double sum = 0;
  for (int t = 0; t < nt; t++) sum += acc[t].e;
  if (want_Hb && H && b) {
    std::fill(H, H + 36, 0.0);
    std::fill(b, b + 6, 0.0);
    for (int t = 0; t < nt; t++) {
      for (int k = 0; k < 36; k++) H[k] += acc[t].H[k];
      for (int k = 0; k < 6; k++) b[k] += acc[t].b[k];
    }
  }
  return sum;
}

double FastGICP::linearize(const double T[16], double* H, double* b) {
  linearize_calls++;
  update_correspondences(T);
  return gicp_accumulate(*this, T, H, b, H != nullptr && b != nullptr);
}

double FastGICP::compute_error(const double T[16]) {
  error_calls++;
  return gicp_accumulate(*this, T, nullptr, nullptr, false);
}

// LSQ:82-91
bool FastGICP::is_converged(const double delta[16]) const {
  double r_max = 0, t_max = 0;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double v = delta[r * 4 + c] - (r == c ? 1.0 : 0.0);
      r_max = std::max(r_max, 1.0 / rotation_epsilon * std::fabs(v));
    }
  for (int r = 0; r < 3; r++) t_max = std::max(t_max, 1.0 / transformation_epsilon * std::fabs(delta[r * 4 + 3]));
  return std::max(r_max, t_max) < 1;
}

// LSQ:125-172
bool FastGICP::step_lm(double x0[16], double delta[16]) {
  double H[36], b[6];
  double y0 = linearize(x0, H, b);
  if (lm_lambda < 0.0) {
    double mx = 0;
    for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i * 7]));
    lm_lambda = lm_init_lambda_factor * mx;
  }
  double nu = 2.0;
  for (int i = 0; i < lm_max_iterations; i++) {
    double A[36], nb[6], d[6];
    for (int k = 0; k < 36; k++) A[k] = H[k] + ((k % 7 == 0) ? lm_lambda : 0.0);
    for (int k = 0; k < 6; k++) nb[k] = -b[k];
    ldlt_solve6(A, nb, d);
    identity_d(delta);
    double R[9];
    so3_exp_matrix(d, R);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) delta[r * 4 + c] = R[r * 3 + c];
    for (int r = 0; r < 3; r++) delta[r * 4 + 3] = d[3 + r];
    double xi[16];
    iso_mul(delta, x0, xi);
    double yi = compute_error(xi);
    double den = 0;
    for (int k = 0; k < 6; k++) den += d[k] * (lm_lambda * d[k] - b[k]);
    double rho = (y0 - yi) / den;
    if (rho < 0) {
      if (is_converged(delta)) return true;
      lm_lambda = nu * lm_lambda;
      nu = 2 * nu;
      continue;
    }
    std::memcpy(x0, xi, sizeof(xi));
    lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
    std::memcpy(final_hessian, H, sizeof(H));
    return true;
  }
  return false;
}

// pcl::Registration::align shell + FG:103-112 + LSQ:53-79
void FastGICP::align(const float* guess, std::vector<P4>* output) {
  linearize_calls = error_calls = 0;
  converged = false;
  for (int i = 0; i < 16; i++) final_transformation[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  if (source->covs.size() != source->pts.size() * 9) calculate_covariances(*source);
  if (target->covs.size() != target->pts.size() * 9) calculate_covariances(*target);

  double x0[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) x0[r * 4 + c] = static_cast<double>(guess[c * 4 + r]);
  lm_lambda = -1.0;
  nr_iterations = 0;
  for (int i = 0; i < max_iterations && !converged; i++) {
    nr_iterations = i;
    double delta[16];
    if (!step_lm(x0, delta)) {
      std::cerr << "lm not converged!!" << std::endl;
      break;
    }
    converged = is_converged(delta);
  }
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) final_transformation[c * 4 + r] = static_cast<float>(x0[r * 4 + c]);
  if (output) {
    output->resize(source->pts.size());
    for (size_t i = 0; i < source->pts.size(); i++) (*output)[i] = transform_point(final_transformation, source->pts[i]);
  }
}

double FastGICP::getFitnessScore(double max_range) {
  return fitness_score(target->tree, source->pts.data(), source->pts.size(), final_transformation, max_range, num_threads);
}

}  // namespace lgs_oracle
