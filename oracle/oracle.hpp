// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Dependency-free CPU restatement of the LiDAR front-end hot path of
// RyuYamamoto/lidar_graph_slam.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library.
//
// PARITY UNPINNED: the reference cannot be compiled in the build container (PCL, Eigen,
// FLANN absent: SURVEY.md section 0.4) and its own tests hold no bit-level golden vectors for
// this path, so the restatement is pinned only by (a) the bundled Velodyne pair converging
// to thirdparty/fast_gicp/data/relative.txt inside the gtest band (0.05 m / 1 deg,
// gicp_test.cpp:147-201) and (b) the README fitness anchors (loose).  See DESIGN.md.
//
// Abbreviations for citations (all under /root/reference):
//   PPF  = points_prefiltering/src/points_prefiltering.cpp
//   NDT  = thirdparty/ndt_omp/include/pclomp/ndt_omp_impl.hpp         NDT.h = .../ndt_omp.h
//   VGC  = thirdparty/ndt_omp/include/pclomp/voxel_grid_covariance_omp_impl.hpp
//   FG   = thirdparty/fast_gicp/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp
//   LSQ  = thirdparty/fast_gicp/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp
//   PCL  = un-vendored PCL 1.12 behaviour restated from its published sources
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <vector>

namespace lgs_oracle {

struct P4 {
  float x, y, z, w;  // w carries intensity
};

inline float sum3f(float a, float b, float c) { return a + (b + c); }   // Eigen unrolled 3-redux order
inline double sum3d(double a, double b, double c) { return a + (b + c); }

// pcl::transformPointCloud(Matrix4f) for PCL >= 1.10: c0*x + (c1*y + (c2*z + c3)).  T is column-major.
inline P4 transform_point(const float* T, const P4& p) {
  P4 r;
  r.x = T[0] * p.x + (T[4] * p.y + (T[8] * p.z + T[12]));
  r.y = T[1] * p.x + (T[5] * p.y + (T[9] * p.z + T[13]));
  r.z = T[2] * p.x + (T[6] * p.y + (T[10] * p.z + T[14]));
  r.w = p.w;
  return r;
}

// ---------------------------------------------------------------------------------------------
// Prefilter: PPF:102-112 (distance_filter), PPF:89-100 (crop), PPF:114-121 -> pcl::VoxelGrid::filter
enum { VG_OK = 0, VG_REFUSED_OVERFLOW = 1 };

struct VoxelGridResult {
  int status = VG_OK;
  std::vector<P4> out;               // centroids, ascending voxel idx
  std::vector<int32_t> out_idx;      // voxel idx of each centroid
  std::vector<int32_t> out_count;    // points per output voxel
  std::vector<int32_t> voxel_idx;    // per INPUT point: voxel idx, -1 if cropped
  std::vector<int32_t> member_rank;  // per INPUT point: row of `out` it was averaged into, -1 if none
  int32_t min_b[3] = {0, 0, 0}, max_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0};
  size_t n_kept = 0;                 // points surviving the crop
};

// range_min < 0 disables the range test; box == nullptr disables the box test.
void prefilter_voxel_grid(const P4* pts, size_t n, const float leaf[3], int min_points_per_voxel,
                          double range_min, const double* box6, VoxelGridResult* res);

// ---------------------------------------------------------------------------------------------
// pclomp::VoxelGridCovariance (VGC:48-442, Leaf = voxel_grid_covariance_omp.h:98-193)
struct Leaf {
  int nr_points = 0;
  double mean[3] = {0, 0, 0};
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // row-major
  double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double evals[3] = {0, 0, 0};
  double evecs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

struct VoxelGridCovariance {
  float leaf_size[3] = {1, 1, 1}, inv_leaf[3] = {1, 1, 1};
  int min_points_per_voxel = 6;
  double min_covar_eigvalue_mult = 0.01;
  int32_t min_b[3] = {0, 0, 0}, max_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0}, divb_mul[3] = {0, 0, 0};
  std::map<size_t, Leaf> leaves;
  bool refused = false;

  void set_leaf_size(float lx, float ly, float lz);
  void build(const P4* pts, size_t n);  // applyFilter, VGC:48-370
  // VGC:373-404 with the DIRECT7 / DIRECT1 / DIRECT26 offset tables (VGC:418-442)
  int neighborhood(const P4& pt, int method, const Leaf** out) const;
};

enum { SEARCH_KDTREE = 0, SEARCH_DIRECT26 = 1, SEARCH_DIRECT7 = 2, SEARCH_DIRECT1 = 3 };  // ndt_omp.h enum order

// ---------------------------------------------------------------------------------------------
// Exact k-NN standing in for pcl::search::KdTree -> FLANN KDTreeSingleIndex<L2_Simple<float>> (PCL).
// Distances ((dx*dx)+(dy*dy))+(dz*dz) in f32; ties broken towards the smaller index so results
// are deterministic (FLANN leaves ties arbitrary).
struct KdTree {
  struct Node {
    int32_t left, right;   // children, or -1
    int32_t begin, end;    // leaf range in `order`
    int32_t dim;
    float split_lo, split_hi;
    float bmin[3], bmax[3];
  };
  std::vector<Node> nodes;
  std::vector<int32_t> order;
  std::vector<P4> pts;  // reordered copy
  void build(const P4* p, size_t n);
  // fills idx/d2 (ascending d2) with min(k, n) entries, returns the count
  int knn(const P4& q, int k, int32_t* idx, float* d2) const;
  size_t size() const { return pts.size(); }

 private:
  int build_rec(int begin, int end);
};

// pcl::Registration::getFitnessScore (PCL; called GBS:321)
double fitness_score(const KdTree& target_tree, const P4* src, size_t n, const float* T_colmajor, double max_range,
                     int num_threads);

// ---------------------------------------------------------------------------------------------
// pclomp::NormalDistributionsTransform (NDT:47-982, NDT.h:71-502)
struct NdtStats {
  int derivative_evals = 0;   // computeDerivatives calls
  int line_search_trials = 0; // inner MT iterations
  int hessian_recomputes = 0; // computeHessian calls
};

struct NDT {
  // parameters (defaults NDT:49-51,71-75)
  float resolution = 1.0f;
  double step_size = 0.1;
  double outlier_ratio = 0.55;
  double transformation_epsilon = 0.1;
  int max_iterations = 35;
  int search_method = SEARCH_DIRECT7;
  int num_threads = 1;

  // state
  std::vector<P4> target, source;
  bool have_source = false;
  VoxelGridCovariance cells;
  KdTree target_tree;       // pcl::Registration::tree_, for getFitnessScore
  bool tree_dirty = true;
  float final_transformation[16];  // column-major
  int nr_iterations = 0;
  bool converged = false;
  double trans_probability = 0;
  double gauss_d1 = 0, gauss_d2 = 0, gauss_d3 = 0;
  NdtStats stats;

  // angular derivative tables (NDT:288-394): double 3-vectors and float 8x4 / 16x4 matrices
  double j_ang_d[8][3];
  double h_ang_d[15][3];
  float j_ang[8][4];
  float h_ang[16][4];

  NDT();
  void setInputTarget(const P4* p, size_t n);   // NDT.h:122-127
  void setInputSource(const P4* p, size_t n);
  void setResolution(float r);                  // NDT.h:132-142
  void align(const float* guess_colmajor, std::vector<P4>* output);  // PCL align shell + NDT:80-171
  double getFitnessScore(double max_range);
  double calculateScore(const std::vector<P4>& trans_cloud) const;   // NDT:934-982

  void compute_gauss();                                              // NDT:86-93
  void computeAngleDerivatives(const double p[6]);                   // NDT:288-394
  double computeDerivatives(double g[6], double H[36], const std::vector<P4>& trans_cloud, const double p[6],
                            bool compute_hessian);                   // NDT:179-285
  void computeHessian(double H[36], const std::vector<P4>& trans_cloud);   // NDT:539-609
  double computeStepLengthMT(const double x[6], double step_dir[6], double step_init, double step_max, double step_min,
                             double& score, double g[6], double H[36], std::vector<P4>& trans_cloud);  // NDT:771-931
  void init();
};

void ndt_convert_transform(const double x[6], float* T_colmajor);   // NDT.h:214-231

// ---------------------------------------------------------------------------------------------
// fast_gicp::FastGICP over fast_gicp::LsqRegistration (FG:8-298, LSQ:8-173)
enum { REG_NONE = 0, REG_MIN_EIG = 1, REG_NORMALIZED_MIN_EIG = 2, REG_PLANE = 3, REG_FROBENIUS = 4 };  // gicp_settings.hpp order

struct Cloud {
  std::vector<P4> pts;
  KdTree tree;
  std::vector<double> covs;  // 9 doubles per point (3x3 block of the reference's Matrix4d), row-major
};

struct FastGICP {
  int num_threads = 1;
  int k_correspondences = 20;
  double corr_dist_threshold = 3.4028234663852886e38;  // FLT_MAX (FG:18)
  int regularization = REG_PLANE;
  int max_iterations = 64;            // LSQ:11
  double rotation_epsilon = 2e-3;     // LSQ:12
  double transformation_epsilon = 5e-4;
  int lm_max_iterations = 10;
  double lm_init_lambda_factor = 1e-9;
  double lm_lambda = -1.0;

  std::shared_ptr<Cloud> source, target;
  std::vector<int32_t> correspondences;
  std::vector<float> sq_distances;
  std::vector<double> mahalanobis;   // 9 doubles per source point
  double final_hessian[36];
  float final_transformation[16];
  int nr_iterations = 0;
  bool converged = false;
  int linearize_calls = 0, error_calls = 0;

  FastGICP();
  void setInputSource(const P4* p, size_t n);
  void setInputTarget(const P4* p, size_t n);
  void swapSourceAndTarget();            // FG:50-57
  void clearSource();
  void clearTarget();
  void align(const float* guess_colmajor, std::vector<P4>* output);
  double getFitnessScore(double max_range);

  void calculate_covariances(Cloud& c);                                  // FG:241-298
  void update_correspondences(const double T[16]);                       // FG:115-152 (T row-major 4x4)
  double linearize(const double T[16], double* H, double* b);            // FG:155-211
  double compute_error(const double T[16]);                              // FG:214-237
  bool step_lm(double x0[16], double delta[16]);                         // LSQ:125-172
  bool is_converged(const double delta[16]) const;                       // LSQ:82-91
};


// ---------------------------------------------------------------------------------------------
// pclomp::GeneralizedIterativeClosestPoint (GO = thirdparty/ndt_omp/include/pclomp/gicp_omp_impl.hpp,
// GO.h = .../gicp_omp.h) over pcl::IterativeClosestPoint / pcl::Registration::align, optimised with PCL's BFGS
// (pcl/registration/bfgs.h: un-vendored, a port of GSL's vector_bfgs2 restated from its published algorithm;
// PARITY UNPINNED like the rest of the PCL internals).
struct BfgsParameters {   // defaults of pcl BFGS::Parameters, then the overrides of GO:212-217
  int max_iters = 400, bracket_iters = 100, section_iters = 100;
  double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5, step_size = 1;
  int order = 3;
};
enum { BFGS_NEGATIVE_GRADIENT_EPSILON = -3, BFGS_NOT_STARTED = -2, BFGS_RUNNING = -1, BFGS_SUCCESS = 0, BFGS_NO_PROGRESS = 1 };

struct PclGICP {
  // GO.h:116-126
  int k_correspondences = 20;
  double gicp_epsilon = 0.001;
  double rotation_epsilon = 2e-3;
  int max_inner_iterations = 20;
  int max_iterations = 200;
  double transformation_epsilon = 5e-4;
  double corr_dist_threshold = 5.0;
  int num_threads = 1;

  std::vector<P4> source, target;
  KdTree source_tree, target_tree;        // tree_reciprocal_ / tree_
  std::vector<double> source_covs, target_covs;   // 9 doubles per point, row-major (GO:48-122)
  std::vector<float> mahalanobis;          // 9 floats per source point: the 3x3 block of mahalanobis_[i] (GO:439-452)
  std::vector<P4> output;                  // the source transformed by the guess (GO:398), w = 1
  std::vector<int32_t> corr_src, corr_tgt; // the sorted correspondence list of the current outer iteration (GO:458-474)
  float base_transformation[16];           // column-major; identity (GO:394)
  float transformation[16], previous_transformation[16], final_transformation[16];
  int nr_iterations = 0;
  bool converged = false;
  int f_calls = 0, df_calls = 0, fdf_calls = 0, inner_iterations_total = 0;

  PclGICP();
  void setInputSource(const P4* p, size_t n);   // GO.h:139-156
  void setInputTarget(const P4* p, size_t n);   // GO.h:164-169
  void align(const float* guess_colmajor, std::vector<P4>* out);   // pcl::Registration::align + GO:370-516
  double getFitnessScore(double max_range);

  void computeCovariances(const std::vector<P4>& cloud, const KdTree& tree, std::vector<double>& covs) const;  // GO:48-122
  void update_correspondences(const float* guess_colmajor);     // GO:404-474
  void applyState(float* t_colmajor, const double x[6]) const;  // GO:518-529
  void computeRDerivative(const double x[6], const double R[9], double g[6]) const;  // GO:125-178
  double functor_f(const double x[6]);                          // GO:245-275
  void functor_df(const double x[6], double g[6]);              // GO:278-330
  void functor_fdf(const double x[6], double& f, double g[6]);  // GO:333-367
  // GO:180-242; returns false when the reference would throw (fewer than 4 correspondences / solver failure)
  bool estimateRigidTransformationBFGS(float* transformation_colmajor);
};


// ---------------------------------------------------------------------------------------------
// pcl::IterativeClosestPoint<PointXYZI, PointXYZI> as GBS:142-151 configures it (the DEFAULT loop-closure method,
// graph_based_slam.param.yaml:9): un-vendored PCL 1.12 (registration/impl/icp.hpp, correspondence_estimation.hpp,
// transformation_estimation_svd.hpp -> pcl::umeyama, default_convergence_criteria.hpp), restated from the published
// sources.  PARITY UNPINNED, and one deliberate deviation: Eigen's f32 column sums and its 3 x n GEMM inside umeyama
// have no specified summation order, so the means and the cross-covariance are accumulated in f64 here and
// rounded to f32 where Eigen holds f32 values; everything after them (JacobiSVD, R, t) is f32 as in PCL.
enum { ICP_NOT_CONVERGED = 0, ICP_ITERATIONS = 1, ICP_TRANSFORM = 2, ICP_ABS_MSE = 3, ICP_REL_MSE = 4, ICP_NO_CORRESPONDENCES = 5 };

struct PclICP {
  // pcl::Registration / IterativeClosestPoint defaults (registration.h, icp.h)
  int max_iterations = 10;
  double transformation_epsilon = 0.0;
  double transformation_rotation_epsilon = 0.0;
  double euclidean_fitness_epsilon = -1.7976931348623157e308;
  double corr_dist_threshold = 1.3407807929942596e154;   // sqrt(DBL_MAX)
  int min_number_correspondences = 3;
  int num_threads = 1;

  std::vector<P4> source, target;
  KdTree target_tree;
  float final_transformation[16], transformation[16], previous_transformation[16];
  int nr_iterations = 0;
  bool converged = false;
  int convergence_state = ICP_NOT_CONVERGED;
  double last_mse = 0;
  long last_correspondences = 0;
  // DefaultConvergenceCriteria is a member of pcl::IterativeClosestPoint: its previous MSE, state and similar-transforms
  // counter survive from one align to the next (they are initialised in its constructor only)
  double crit_prev_mse = 1.7976931348623157e308;
  int crit_state = ICP_NOT_CONVERGED, crit_similar = 0;

  PclICP();
  void setInputSource(const P4* p, size_t n);
  void setInputTarget(const P4* p, size_t n);
  void align(const float* guess_colmajor, std::vector<P4>* out);
  double getFitnessScore(double max_range);
  // one correspondence + estimation step on `cloud` (the current input_transformed): the 17 sums
  // {count, sum d2, sum p[3], sum q[3], sum q p^T[9]} and the estimated transformation_ (column-major)
  bool estimate_step(const std::vector<P4>& cloud, double sums[17], float* T_colmajor) const;
};

}  // namespace lgs_oracle
