// ORACLE (test infrastructure).  Prefilter (crop + pcl::VoxelGrid) and pclomp::VoxelGridCovariance.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <numeric>

#include "linalg.hpp"
#include "oracle.hpp"

namespace lgs_oracle {

// ---------------------------------------------------------------------------------------------
// PPF:102-112 distance_filter: keep iff min_distance(double) < (double) float-norm.
// PPF:89-100 crop: strict inequalities, double limits against float coordinates.
// Then pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.12 voxel_grid.hpp; index arithmetic is
// restated in-tree at VGC:67-103,218-223):
//   bbox (getMinMax3D), overflow refusal, min_b/div_b/divb_mul, per-point key, sort by key,
//   one CentroidPoint (f32 accumulators, all fields) per distinct key, ascending key order.
// The reference's sort is unstable (boost spreadsort / std::sort), so the within-voxel f32
// summation order is unspecified there; the oracle fixes it to ascending point index.
void prefilter_voxel_grid(const P4* pts, size_t n, const float leaf[3], int min_points_per_voxel, double range_min,
                          const double* box, VoxelGridResult* res) {
  res->status = VG_OK;
  res->out.clear();
  res->out_idx.clear();
  res->out_count.clear();
  res->voxel_idx.assign(n, -1);
  res->member_rank.assign(n, -1);

  std::vector<uint32_t> kept;
  kept.reserve(n);
  for (size_t i = 0; i < n; i++) {
    const P4& p = pts[i];
    // pcl::VoxelGrid::applyFilter / getMinMax3D on a cloud that is not is_dense (restated at VGC:210-215): points with a
    // non-finite coordinate are skipped.  Every cloud is treated as not dense here (a dense cloud has none to skip).
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    if (range_min >= 0.0) {
      const float norm = std::sqrt(sum3f(p.x * p.x, p.y * p.y, p.z * p.z));  // getVector3fMap().norm()
      const double distance = norm;
      if (!(range_min < distance)) continue;
    }
    if (box) {
      if (!((box[0] < p.x && p.x < box[1]) && (box[2] < p.y && p.y < box[3]) && (box[4] < p.z && p.z < box[5]))) continue;
    }
    kept.push_back(uint32_t(i));
  }
  res->n_kept = kept.size();
  if (kept.empty()) return;

  float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};  // VoxelGrid::setLeafSize
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t i : kept) {
    const float c[3] = {pts[i].x, pts[i].y, pts[i].z};
    for (int a = 0; a < 3; a++) {
      mn[a] = std::min(mn[a], c[a]);
      mx[a] = std::max(mx[a], c[a]);
    }
  }
  int64_t d[3];
  for (int a = 0; a < 3; a++) d[a] = static_cast<int64_t>((mx[a] - mn[a]) * inv[a]) + 1;
  if (d[0] * d[1] * d[2] > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    // pcl::VoxelGrid warns and copies the input to the output
    res->status = VG_REFUSED_OVERFLOW;
    for (uint32_t i : kept) res->out.push_back(pts[i]);
    return;
  }
  for (int a = 0; a < 3; a++) {
    res->min_b[a] = static_cast<int>(std::floor(mn[a] * inv[a]));
    res->max_b[a] = static_cast<int>(std::floor(mx[a] * inv[a]));
    res->div_b[a] = res->max_b[a] - res->min_b[a] + 1;
  }
  const int mul[3] = {1, res->div_b[0], res->div_b[0] * res->div_b[1]};

  std::vector<std::pair<uint32_t, uint32_t>> iv;  // (idx, point index)
  iv.reserve(kept.size());
  for (uint32_t i : kept) {
    int ijk0 = static_cast<int>(std::floor(pts[i].x * inv[0]) - static_cast<float>(res->min_b[0]));
    int ijk1 = static_cast<int>(std::floor(pts[i].y * inv[1]) - static_cast<float>(res->min_b[1]));
    int ijk2 = static_cast<int>(std::floor(pts[i].z * inv[2]) - static_cast<float>(res->min_b[2]));
    int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    res->voxel_idx[i] = idx;
    iv.emplace_back(static_cast<uint32_t>(idx), i);
  }
  std::stable_sort(iv.begin(), iv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });

  size_t index = 0;
  while (index < iv.size()) {
    size_t j = index + 1;
    while (j < iv.size() && iv[j].first == iv[index].first) ++j;
    if (j - index >= static_cast<size_t>(std::max(min_points_per_voxel, 0))) {
      float sx = 0, sy = 0, sz = 0, si = 0;  // CentroidPoint<PointXYZI>: AccumulatorXYZ + AccumulatorIntensity
      for (size_t li = index; li < j; ++li) {
        const P4& p = pts[iv[li].second];
        sx += p.x;
        sy += p.y;
        sz += p.z;
        si += p.w;
        res->member_rank[iv[li].second] = static_cast<int32_t>(res->out.size());
      }
      const float cnt = static_cast<float>(j - index);
      res->out.push_back(P4{sx / cnt, sy / cnt, sz / cnt, si / cnt});
      res->out_idx.push_back(static_cast<int32_t>(iv[index].first));
      res->out_count.push_back(static_cast<int32_t>(j - index));
    }
    index = j;
  }
}

// ---------------------------------------------------------------------------------------------
void VoxelGridCovariance::set_leaf_size(float lx, float ly, float lz) {
  leaf_size[0] = lx;
  leaf_size[1] = ly;
  leaf_size[2] = lz;
  for (int a = 0; a < 3; a++) inv_leaf[a] = 1.0f / leaf_size[a];
}

// VGC:48-370.  is_dense input, no filter field, downsample_all_data irrelevant to NDT (centroid
// cloud is only used by the unused kd-tree, VGC.h:298-302).
void VoxelGridCovariance::build(const P4* pts, size_t n) {
  leaves.clear();
  refused = false;
  if (n == 0) {
    refused = true;
    return;
  }
  // VGC:210-215 (and getMinMax3D at VGC:67): a cloud that is not is_dense has its non-finite points skipped
  auto finite = [&](size_t i) { return std::isfinite(pts[i].x) && std::isfinite(pts[i].y) && std::isfinite(pts[i].z); };
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  size_t n_finite = 0;
  for (size_t i = 0; i < n; i++) {
    if (!finite(i)) continue;
    n_finite++;
    const float c[3] = {pts[i].x, pts[i].y, pts[i].z};
    for (int a = 0; a < 3; a++) {
      mn[a] = std::min(mn[a], c[a]);
      mx[a] = std::max(mx[a], c[a]);
    }
  }
  if (n_finite == 0) {
    refused = true;
    return;
  }
  // VGC:75-84
  int64_t d[3];
  for (int a = 0; a < 3; a++) d[a] = static_cast<int64_t>((mx[a] - mn[a]) * inv_leaf[a]) + 1;
  if (d[0] * d[1] * d[2] > std::numeric_limits<int32_t>::max()) {
    refused = true;
    return;
  }
  // VGC:87-103
  for (int a = 0; a < 3; a++) {
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv_leaf[a]));
    max_b[a] = static_cast<int>(std::floor(mx[a] * inv_leaf[a]));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  divb_mul[0] = 1;
  divb_mul[1] = div_b[0];
  divb_mul[2] = div_b[0] * div_b[1];

  // first pass, VGC:209-263
  for (size_t cp = 0; cp < n; ++cp) {
    if (!finite(cp)) continue;
    int ijk0 = static_cast<int>(std::floor(pts[cp].x * inv_leaf[0]) - static_cast<float>(min_b[0]));
    int ijk1 = static_cast<int>(std::floor(pts[cp].y * inv_leaf[1]) - static_cast<float>(min_b[1]));
    int ijk2 = static_cast<int>(std::floor(pts[cp].z * inv_leaf[2]) - static_cast<float>(min_b[2]));
    int idx = ijk0 * divb_mul[0] + ijk1 * divb_mul[1] + ijk2 * divb_mul[2];
    Leaf& leaf = leaves[static_cast<size_t>(idx)];
    const double p[3] = {pts[cp].x, pts[cp].y, pts[cp].z};
    for (int a = 0; a < 3; a++) leaf.mean[a] += p[a];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) leaf.cov[a * 3 + b] += p[a] * p[b];
    ++leaf.nr_points;
  }

  // second pass, VGC:282-367
  for (auto& kv : leaves) {
    Leaf& leaf = kv.second;
    double pt_sum[3] = {leaf.mean[0], leaf.mean[1], leaf.mean[2]};
    for (int a = 0; a < 3; a++) leaf.mean[a] /= leaf.nr_points;
    if (leaf.nr_points < min_points_per_voxel) continue;

    const double npts = leaf.nr_points;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
        leaf.cov[a * 3 + b] = (leaf.cov[a * 3 + b] - 2 * (pt_sum[a] * leaf.mean[b])) / npts + leaf.mean[a] * leaf.mean[b];
    const double f = (leaf.nr_points - 1.0) / leaf.nr_points;
    for (int k = 0; k < 9; k++) leaf.cov[k] *= f;

    double ev[3], V[9];
    self_adjoint_eigen3(leaf.cov, ev, V);
    std::copy(V, V + 9, leaf.evecs);
    if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
      leaf.nr_points = -1;
      continue;
    }
    const double min_covar_eigvalue = min_covar_eigvalue_mult * ev[2];
    if (ev[0] < min_covar_eigvalue) {
      ev[0] = min_covar_eigvalue;
      if (ev[1] < min_covar_eigvalue) ev[1] = min_covar_eigvalue;
      // cov = evecs * diag(ev) * evecs^-1   (VGC:355; diag held as a full 3x3)
      double D[9] = {ev[0], 0, 0, 0, ev[1], 0, 0, 0, ev[2]};
      double Vinv[9], VD[9];
      inverse3(V, Vinv);
      matmul3(V, D, VD);
      matmul3(VD, Vinv, leaf.cov);
    }
    for (int a = 0; a < 3; a++) leaf.evals[a] = ev[a];
    inverse3(leaf.cov, leaf.icov);
    double mxc = leaf.icov[0], mnc = leaf.icov[0];
    for (int k = 1; k < 9; k++) {
      mxc = std::max(mxc, leaf.icov[k]);
      mnc = std::min(mnc, leaf.icov[k]);
    }
    if (mxc == std::numeric_limits<float>::infinity() || mnc == -std::numeric_limits<float>::infinity()) leaf.nr_points = -1;
  }
}

// VGC:373-404; offsets VGC:418-442 and pcl::getAllNeighborCellIndices (PCL) for DIRECT26.
int VoxelGridCovariance::neighborhood(const P4& pt, int method, const Leaf** out) const {
  if (refused) return 0;
  static const int off7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  int ijk[3] = {static_cast<int>(std::floor(pt.x / leaf_size[0])), static_cast<int>(std::floor(pt.y / leaf_size[1])),
                static_cast<int>(std::floor(pt.z / leaf_size[2]))};
  int cnt = 0;
  auto probe = [&](int dx, int dy, int dz) {
    const int dd[3] = {dx, dy, dz};
    for (int a = 0; a < 3; a++) {
      if (!(min_b[a] - ijk[a] <= dd[a] && max_b[a] - ijk[a] >= dd[a])) return;
    }
    int key = 0;
    for (int a = 0; a < 3; a++) key += (ijk[a] + dd[a] - min_b[a]) * divb_mul[a];
    auto it = leaves.find(static_cast<size_t>(key));
    if (it != leaves.end() && it->second.nr_points >= min_points_per_voxel) out[cnt++] = &it->second;
  };
  if (method == SEARCH_DIRECT1) {
    probe(0, 0, 0);
  } else if (method == SEARCH_DIRECT26) {
    // pcl::getAllNeighborCellIndices() (PCL voxel_grid.h): the 13 "half" offsets -- (i,j,-1) for
    // i,j in -1..1, then (i,-1,0) for i in -1..1, then (-1,0,0) -- followed by their negations.
    // 26 cells: the centre cell itself is NOT visited.
    int half[13][3];
    int c = 0;
    for (int i = -1; i < 2; i++)
      for (int j = -1; j < 2; j++) {
        half[c][0] = i; half[c][1] = j; half[c][2] = -1; c++;
      }
    for (int i = -1; i < 2; i++) {
      half[c][0] = i; half[c][1] = -1; half[c][2] = 0; c++;
    }
    half[c][0] = -1; half[c][1] = 0; half[c][2] = 0;
    for (int t = 0; t < 13; t++) probe(half[t][0], half[t][1], half[t][2]);
    for (int t = 0; t < 13; t++) probe(-half[t][0], -half[t][1], -half[t][2]);
  } else {
    for (int t = 0; t < 7; t++) probe(off7[t][0], off7[t][1], off7[t][2]);
  }
  return cnt;
}

}  // namespace lgs_oracle
