// Host build of the NDT optimiser state machine (csrc/ndt_opt.cuh: the code ndt_align_kernel runs on the device and the
// host-stepped driver runs on the CPU), behind a tiny C interface so that a CPU test can drive it with the ORACLE's
// derivative evaluations in place of the CUDA kernels: Newton step, More-Thuente search, pose -> transform, restated
// sinf / cosf, exit rules - everything between two evaluations - against the oracle's own optimiser, without a GPU.
//   g++ -O2 -std=c++17 -shared -fPIC -I lidar_graph_slam_b200/csrc tests/ndt_machine_host.cpp -o libndt_machine_host.so
#include <cstring>

#include "ndt_opt.cuh"

using namespace lgs::ndtopt;

struct Host {
  Machine m;
  Command c;
  bool first = true;
};

extern "C" {

void* mh_create() { return new Host(); }
void mh_destroy(void* h) { delete static_cast<Host*>(h); }

// NDT:95-119: the pose of the guess and the first command
void mh_begin(void* hp, const float* guess16, double step_size, double trans_eps, int max_iter, double n_in, int exact_solve) {
  Host* h = static_cast<Host*>(hp);
  double p0[6];
  matrix_to_pose(guess16, p0);
  h->m.begin(p0, guess16, step_size, trans_eps, max_iter, n_in, &h->c, exact_solve);
  h->first = true;
}

// the evaluation to run: column-major transform, the pose it belongs to (for computeAngleDerivatives), the mode
void mh_command(void* hp, float* T16, double* pose6, int* mode) {
  Host* h = static_cast<Host*>(hp);
  memcpy(T16, h->c.T, sizeof(float) * 16);
  memcpy(pose6, h->first ? h->m.p : h->m.x_t, sizeof(double) * 6);
  *mode = h->c.mode;
}

// sums: the row an evaluation kernel leaves (score, g[6], upper triangle, strict lower triangle; mode 2: upper triangle)
int mh_advance(void* hp, const double* sums) {
  Host* h = static_cast<Host*>(hp);
  h->first = false;
  return h->m.advance(sums, &h->c) ? 1 : 0;
}

void mh_result(void* hp, float* T16, int* counts5, double* trans_probability) {
  Host* h = static_cast<Host*>(hp);
  memcpy(T16, h->m.final_T, sizeof(float) * 16);
  counts5[0] = h->m.nr_iterations;
  counts5[1] = h->m.converged;
  counts5[2] = h->m.evals;
  counts5[3] = h->m.trials;
  counts5[4] = h->m.hess_recomputes;
  *trans_probability = h->m.trans_probability;
}
}
