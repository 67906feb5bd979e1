// ORACLE (test infrastructure, never shipped, never on the product path).
//
// pcl::StatisticalOutlierRemoval<PointXYZI>::filter as the prefilter node calls it right after the voxel grid
// (points_prefiltering/src/points_prefiltering.cpp:79-80,132-140; defaults mean_k = 30, stddev = 1.2 in
// launch/points_prefiltering.launch.xml:4-5).  The algorithm lives in PCL (un-vendored; ROS 2 Humble => 1.12.1,
// filters/impl/statistical_outlier_removal.hpp applyFilterIndices) and is restated here from its published form:
//
//   for every point i: exact (mean_k + 1)-NN of the point in the cloud itself (entry 0 is the query point);
//       dist_sum (double) = sum_{k=1..mean_k} sqrt(d2_k)  (f32 square root of the f32 squared distance);
//       distances[i] = float(dist_sum / mean_k)
//   sum, sq_sum (double) accumulated serially in index order over the float distances (the square is an f32 product);
//   mean = sum / n;  variance = (sq_sum - sum * sum / n) / (n - 1);  threshold = mean + std_mul * sqrt(variance)
//   keep point i  iff  !(distances[i] > threshold)        (setNegative(true) keeps the complement)
//
// Parity unpinned: no reference test covers this filter; PCL's internals are restated, not linked.  A cloud with
// fewer than mean_k + 1 points makes PCL read past the k-NN result (undefined); here the available neighbours are
// summed and still divided by mean_k, and the CUDA path does the same.
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.hpp"

namespace lgs_oracle {

struct OutlierResult {
  std::vector<float> distances;
  std::vector<unsigned char> keep;
  std::vector<P4> points;
  double mean = 0, stddev = 0, threshold = 0;
};

static OutlierResult* outlier_filter(const P4* pts, long n, int mean_k, double std_mul, int negative) {
  OutlierResult* r = new OutlierResult;
  r->distances.assign(n, 0.f);
  r->keep.assign(n, 0);
  if (n == 0) return r;
  KdTree tree;
  tree.build(pts, n);
  const int k = mean_k + 1;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) {
    std::vector<int32_t> idx(k);
    std::vector<float> d2(k);
    const int found = tree.knn(pts[i], k, idx.data(), d2.data());
    double dist_sum = 0.0;
    for (int j = 1; j < found; j++) dist_sum += static_cast<double>(std::sqrt(d2[j]));
    r->distances[i] = static_cast<float>(dist_sum / mean_k);
  }
  double sum = 0, sq_sum = 0;
  for (long i = 0; i < n; i++) {
    const float d = r->distances[i];
    sum += d;
    sq_sum += d * d;  // f32 product, widened
  }
  const double valid = static_cast<double>(n);
  r->mean = sum / valid;
  const double variance = (sq_sum - sum * sum / valid) / (valid - 1);
  r->stddev = std::sqrt(variance);
  r->threshold = r->mean + std_mul * r->stddev;
  for (long i = 0; i < n; i++) {
    const bool out = r->distances[i] > r->threshold;
    const bool kept = negative ? out : !out;
    r->keep[i] = kept ? 1 : 0;
    if (kept) r->points.push_back(pts[i]);
  }
  return r;
}

}  // namespace lgs_oracle

using namespace lgs_oracle;

extern "C" {
void* orc_sor_run(const float* pts, long n, int mean_k, double std_mul, int negative) {
  return outlier_filter(reinterpret_cast<const P4*>(pts), n, mean_k, std_mul, negative);
}
long orc_sor_out_n(void* h) { return static_cast<long>(static_cast<OutlierResult*>(h)->points.size()); }
void orc_sor_get(void* h, float* out_pts, float* distances, unsigned char* keep, double* stats3) {
  OutlierResult* r = static_cast<OutlierResult*>(h);
  if (out_pts && !r->points.empty()) std::memcpy(out_pts, r->points.data(), r->points.size() * sizeof(P4));
  if (distances && !r->distances.empty()) std::memcpy(distances, r->distances.data(), r->distances.size() * sizeof(float));
  if (keep && !r->keep.empty()) std::memcpy(keep, r->keep.data(), r->keep.size());
  if (stats3) {
    stats3[0] = r->mean;
    stats3[1] = r->stddev;
    stats3[2] = r->threshold;
  }
}
void orc_sor_free(void* h) { delete static_cast<OutlierResult*>(h); }
}
