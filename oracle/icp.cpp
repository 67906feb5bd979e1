// ORACLE (test infrastructure).  pcl::IterativeClosestPoint as the graph SLAM node configures it (GBS:142-151),
// restated from PCL 1.12's published sources (un-vendored; PARITY UNPINNED):
//   icp.hpp computeTransformation                      -> PclICP::align
//   correspondence_estimation.hpp determineCorrespondences (1-NN, kept unless d2 > max_dist^2)
//   transformation_estimation_svd.hpp -> pcl::umeyama (no scaling)   -> umeyama_from_sums
//   default_convergence_criteria.hpp hasConverged      -> Criteria
// Deviation (see oracle.hpp): the means and the cross-covariance are accumulated in f64 and rounded to f32.
#include <cmath>
#include <cstring>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "linalg.hpp"
#include "oracle.hpp"

namespace lgs_oracle {

namespace {

void identity_f(float* T) {
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}
void mul4f(const float* A, const float* B, float* C) {
  float R[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) R[c * 4 + r] = ((A[0 * 4 + r] * B[c * 4 + 0] + A[1 * 4 + r] * B[c * 4 + 1]) + A[2 * 4 + r] * B[c * 4 + 2]) + A[3 * 4 + r] * B[c * 4 + 3];
  std::memcpy(C, R, sizeof(R));
}

float det3f(const float* m) {  // Eigen 3x3 determinant: cofactor expansion along the first column
  auto h = [&](int a, int b, int c) { return m[0 * 3 + a] * (m[1 * 3 + b] * m[2 * 3 + c] - m[1 * 3 + c] * m[2 * 3 + b]); };
  return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
}

// pcl::umeyama(src, dst, with_scaling = false) from the sums of one correspondence set
void umeyama_from_sums(const double* s, float* T) {
  const double n = s[0];
  float src_mean[3], dst_mean[3], sigma[9];
  for (int a = 0; a < 3; a++) {
    src_mean[a] = static_cast<float>(s[2 + a] / n);
    dst_mean[a] = static_cast<float>(s[5 + a] / n);
  }
  // sigma = (1/n) sum (q - q_mean)(p - p_mean)^T = (sum q p^T) / n - q_mean p_mean^T, f64 then rounded
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) sigma[r * 3 + c] = static_cast<float>(s[8 + r * 3 + c] / n - (s[5 + r] / n) * (s[2 + c] / n));
  float U[9], S[3], V[9];
  jacobi_svd<3, float>(sigma, U, S, V);
  float sgn[3] = {1.f, 1.f, 1.f};
  if (det3f(U) * det3f(V) < 0) sgn[2] = -1.f;
  float R[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[r * 3 + c] = ((U[r * 3 + 0] * sgn[0]) * V[c * 3 + 0] + (U[r * 3 + 1] * sgn[1]) * V[c * 3 + 1]) + (U[r * 3 + 2] * sgn[2]) * V[c * 3 + 2];
  identity_f(T);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[c * 4 + r] = R[r * 3 + c];
  for (int r = 0; r < 3; r++) T[12 + r] = dst_mean[r] - ((R[r * 3 + 0] * src_mean[0] + R[r * 3 + 1] * src_mean[1]) + R[r * 3 + 2] * src_mean[2]);
}

// pcl::registration::DefaultConvergenceCriteria
struct Criteria {
  int max_iterations = 1000;
  int max_iterations_similar_transforms = 0, iterations_similar_transforms = 0;
  double rotation_threshold = 0.99999, translation_threshold = 3e-4 * 3e-4;
  double mse_threshold_relative = 0.00001, mse_threshold_absolute = 1e-12;
  double prev_mse = std::numeric_limits<double>::max(), cur_mse = std::numeric_limits<double>::max();
  int state = ICP_NOT_CONVERGED;
  bool has_converged(int iterations, const float* T, double mse) {
    if (state != ICP_NOT_CONVERGED) {
      iterations_similar_transforms = 0;
      state = ICP_NOT_CONVERGED;
    }
    bool is_similar = false;
    if (iterations >= max_iterations) {
      state = ICP_ITERATIONS;
      return true;
    }
    const double cos_angle = 0.5 * (T[0] + T[5] + T[10] - 1);  // f32 sum widened by the 0.5 *
    const double translation_sqr = T[12] * T[12] + T[13] * T[13] + T[14] * T[14];
    if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) {
      if (iterations_similar_transforms >= max_iterations_similar_transforms) {
        state = ICP_TRANSFORM;
        return true;
      }
      is_similar = true;
    }
    cur_mse = mse;
    if (std::fabs(cur_mse - prev_mse) < mse_threshold_absolute) {
      if (iterations_similar_transforms >= max_iterations_similar_transforms) {
        state = ICP_ABS_MSE;
        return true;
      }
      is_similar = true;
    }
    if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_threshold_relative) {
      if (iterations_similar_transforms >= max_iterations_similar_transforms) {
        state = ICP_REL_MSE;
        return true;
      }
      is_similar = true;
    }
    if (is_similar)
      ++iterations_similar_transforms;
    else
      iterations_similar_transforms = 0;
    prev_mse = cur_mse;
    return false;
  }
};

}  // namespace

PclICP::PclICP() {
#ifdef _OPENMP
  num_threads = omp_get_max_threads();
#endif
  identity_f(final_transformation);
  identity_f(transformation);
  identity_f(previous_transformation);
}
void PclICP::setInputSource(const P4* p, size_t n) { source.assign(p, p + n); }
void PclICP::setInputTarget(const P4* p, size_t n) {
  target.assign(p, p + n);
  target_tree.build(p, n);
}

bool PclICP::estimate_step(const std::vector<P4>& cloud, double sums[17], float* T) const {
  const long n = static_cast<long>(cloud.size());
  const double max_dist_sqr = corr_dist_threshold * corr_dist_threshold;
  std::vector<int32_t> match(n, -1);
  std::vector<float> dist(n, 0.f);
#pragma omp parallel for num_threads(num_threads) schedule(static)
  for (long i = 0; i < n; i++) {
    int32_t id = -1;
    float d2 = 0;
    if (target_tree.knn(cloud[i], 1, &id, &d2) == 0) continue;
    if (static_cast<double>(d2) > max_dist_sqr) continue;
    match[i] = id;
    dist[i] = d2;
  }
  for (int k = 0; k < 17; k++) sums[k] = 0;
  for (long i = 0; i < n; i++) {
    if (match[i] < 0) continue;
    const P4& p = cloud[i];
    const P4& q = target[match[i]];
    sums[0] += 1.0;
    sums[1] += static_cast<double>(dist[i]);
    const double pv[3] = {p.x, p.y, p.z}, qv[3] = {q.x, q.y, q.z};
    for (int a = 0; a < 3; a++) sums[2 + a] += pv[a];
    for (int a = 0; a < 3; a++) sums[5 + a] += qv[a];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) sums[8 + r * 3 + c] += qv[r] * pv[c];
  }
  if (static_cast<int>(sums[0]) < min_number_correspondences) return false;
  umeyama_from_sums(sums, T);
  return true;
}

void PclICP::align(const float* guess, std::vector<P4>* out) {
  nr_iterations = 0;
  converged = false;
  std::memcpy(final_transformation, guess, sizeof(final_transformation));
  std::vector<P4> cloud(source.size());
  bool is_identity = true;
  for (int i = 0; i < 16; i++)
    if (guess[i] != ((i % 5 == 0) ? 1.0f : 0.0f)) is_identity = false;
  if (!is_identity)
    for (size_t i = 0; i < source.size(); i++) cloud[i] = transform_point(guess, source[i]);
  else
    cloud = source;
  identity_f(transformation);
  Criteria crit;
  crit.prev_mse = crit_prev_mse;
  crit.state = crit_state;
  crit.iterations_similar_transforms = crit_similar;
  crit.max_iterations = max_iterations;
  crit.mse_threshold_relative = euclidean_fitness_epsilon;
  crit.translation_threshold = transformation_epsilon;
  crit.rotation_threshold = transformation_rotation_epsilon > 0 ? transformation_rotation_epsilon : 1.0 - transformation_epsilon;
  do {
    std::memcpy(previous_transformation, transformation, sizeof(transformation));
    double sums[17];
    if (!estimate_step(cloud, sums, transformation)) {
      crit.state = ICP_NO_CORRESPONDENCES;
      converged = false;
      break;
    }
    last_correspondences = static_cast<long>(sums[0]);
    last_mse = sums[1] / sums[0];
    for (auto& p : cloud) p = transform_point(transformation, p);
    mul4f(transformation, final_transformation, final_transformation);
    ++nr_iterations;
    converged = crit.has_converged(nr_iterations, transformation, last_mse);
  } while (crit.state == ICP_NOT_CONVERGED);
  convergence_state = crit.state;
  crit_prev_mse = crit.prev_mse;
  crit_state = crit.state;
  crit_similar = crit.iterations_similar_transforms;
  if (out) {
    out->resize(source.size());
    for (size_t i = 0; i < source.size(); i++) (*out)[i] = transform_point(final_transformation, source[i]);
  }
}

double PclICP::getFitnessScore(double max_range) {
  return fitness_score(target_tree, source.data(), source.size(), final_transformation, max_range, num_threads);
}

}  // namespace lgs_oracle
