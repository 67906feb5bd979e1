// Pins the pieces of csrc/ndt_opt.cuh that the device-resident NDT optimiser evaluates in a different FORM than the
// reference writes them, against the reference's form, on the host:
//   * angle_entry (table-driven, one entry per thread on the GPU) == the expressions of NDT:329-392, bit for bit, signed
//     zeros included, over random poses with and without the small-angle snap;
//   * Machine::pose_entry (one transform entry per thread, no exchange) == the staged Eigen-order product of pose_to_matrix;
//   * the block elimination of schur_solve6 == the JacobiSVD solve the reference uses (NDT:127-129) to 1e-8 relative on
//     regular symmetric systems (definite and indefinite), and it refuses singular / non-finite ones (-> JacobiSVD).
//   g++ -O2 -std=c++17 -I lidar_graph_slam_b200/csrc tests/ndt_opt_check.cpp -o ndt_opt_check -lm && ./ndt_opt_check
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "ndt_opt.cuh"

using namespace lgs::ndtopt;

int main() {
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> ang(-3.2, 3.2), small(-2e-4, 2e-4), u(-1.0, 1.0);
  long failures = 0, refused = 0, loose = 0;
  for (int it = 0; it < 20000; it++) {
    double x[6] = {u(rng) * 50, u(rng) * 50, u(rng) * 5, ang(rng), ang(rng), ang(rng)};
    if (it % 3 == 0) x[3 + it % 3] = small(rng);
    if (it % 7 == 0) x[4] = small(rng);
    if (it % 11 == 0) x[3] = x[4] = x[5] = 0.0;
    Trig t;
    trig_of_pose(x, &t);
    double Jd[8][3], Hd[15][3], Jd2[8][3], Hd2[15][3];
    float Jf[8][3], Hf[15][3], Jf2[8][3], Hf2[15][3];
    angle_tables(t, Jd, Hd, Jf, Hf);
    angle_tables_explicit(t, Jd2, Hd2, Jf2, Hf2);
    if (memcmp(Jd, Jd2, sizeof(Jd)) || memcmp(Hd, Hd2, sizeof(Hd)) || memcmp(Jf, Jf2, sizeof(Jf)) || memcmp(Hf, Hf2, sizeof(Hf))) {
      if (failures < 5) printf("angle tables differ at pose %d\n", it);
      failures++;
    }
  }
  // the per-entry form of the transform (Machine::pose_entry, one GPU thread per entry) == convertTransform's staged form
  for (int it = 0; it < 20000; it++) {
    Machine m;
    Command c;
    memset(&m, 0, sizeof(m));
    for (int i = 0; i < 3; i++) m.x_t[i] = u(rng) * 80;
    for (int i = 3; i < 6; i++) m.x_t[i] = (it % 5 == 0) ? small(rng) : ang(rng);
    trig_of_pose(m.x_t, &m.trig);
    float T[16];
    pose_to_matrix(m.x_t, m.trig, T);
    for (int i = 0; i < 16; i++) m.pose_entry(i, &c);
    if (memcmp(T, c.T, sizeof(T)) || memcmp(T, m.final_T, sizeof(T))) {
      if (failures < 5) printf("pose_entry differs at pose %d\n", it);
      failures++;
    }
  }
  // block elimination vs JacobiSVD on systems conditioned like an NDT Hessian (rotations ~1e3 times stiffer), positive and
  // negative definite and indefinite, symmetric up to the f32 rounding of the reference's terms (the two triangles of its
  // Hessian differ by ~1e-8 relative)
  for (int it = 0; it < 3000; it++) {
    double A[36], H[36], g[6];
    for (double& v : A) v = u(rng);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += A[k * 6 + i] * A[k * 6 + j] * ((it % 3 == 2 && k < 2) ? -1.0 : 1.0);
        H[i * 6 + j] = s * (i >= 3 ? 30.0 : 1.0) * (j >= 3 ? 30.0 : 1.0) * (it % 3 == 1 ? -1.0 : 1.0);
      }
    for (int i = 0; i < 6; i++) H[i * 6 + i] += (it % 3 == 1 ? -0.05 : 0.05);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < i; j++) H[i * 6 + j] *= 1.0 + 3e-8 * u(rng);
    for (double& v : g) v = u(rng) * 100;
    double neg_g[6], xs[6], xb[6];
    for (int i = 0; i < 6; i++) neg_g[i] = -g[i];
    lgs::m::svd_solve<6>(H, neg_g, xs);
    if (!schur_solve6(H, neg_g, xb)) {
      refused++;
      continue;
    }
    double nrm = 0, err = 0;
    for (int i = 0; i < 6; i++) {
      nrm += xs[i] * xs[i];
      err += (xs[i] - xb[i]) * (xs[i] - xb[i]);
    }
    // both solutions carry ~cond * 1e-16 of error: the bar is on the residual of the elimination, and on the difference
    // wherever the SVD's own residual says the system is well conditioned
    auto residual = [&](const double* x) {
      double r2 = 0, h2 = 0, x2 = 0;
      for (int i = 0; i < 6; i++) {
        double r = -neg_g[i];
        for (int j = 0; j < 6; j++) r += H[i * 6 + j] * x[j], h2 += H[i * 6 + j] * H[i * 6 + j];
        r2 += r * r;
        x2 += x[i] * x[i];
      }
      return std::sqrt(r2) / (std::sqrt(h2) * std::sqrt(x2));
    };
    if (!(residual(xb) <= 1e-15) || !(std::sqrt(err) <= 1e-9 * std::sqrt(nrm))) {
      if (failures < 5) printf("system %d: relative difference %.3e, residuals %.3e (elimination) %.3e (SVD)\n", it, std::sqrt(err / nrm), residual(xb), residual(xs));
      failures++;
    }
    if (std::sqrt(err) > 1e-11 * std::sqrt(nrm)) loose++;
  }
  if (loose > 5) failures++, printf("%ld of 3000 solutions further than 1e-11 from the SVD's\n", loose);
  if (refused > 300) failures++, printf("block elimination refused %ld of 3000 regular systems\n", refused);
  {  // singular and non-finite systems are handed to the SVD
    double H[36] = {0}, b[6] = {1, 2, 3, 4, 5, 6}, x[6];
    if (schur_solve6(H, b, x)) failures++, printf("zero system accepted\n");
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) H[i * 6 + j] = (i + 1.0) * (j + 1.0);  // rank 1
    if (schur_solve6(H, b, x)) failures++, printf("rank-1 system accepted\n");
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) H[i * 6 + j] = (i == j) ? 2.0 : 0.1;
    H[4 * 6 + 5] = NAN;
    if (schur_solve6(H, b, x)) failures++, printf("NaN system accepted\n");
    H[4 * 6 + 5] = 0.1;
    if (!schur_solve6(H, b, x)) failures++, printf("regular system refused\n");
    // a visibly non-symmetric system: the solution is that of the full matrix, not of a mirrored triangle
    H[3 * 6 + 1] = 0.7;
    double xs[6];
    lgs::m::svd_solve<6>(H, b, xs);
    if (!schur_solve6(H, b, x)) failures++, printf("non-symmetric system refused\n");
    for (int i = 0; i < 6; i++)
      if (!(std::fabs(x[i] - xs[i]) <= 1e-12 * (1.0 + std::fabs(xs[i])))) failures++, printf("non-symmetric system: x[%d] %.17g vs %.17g\n", i, x[i], xs[i]);
  }
  printf("%ld failures\n", failures);
  return failures ? 1 : 0;
}
