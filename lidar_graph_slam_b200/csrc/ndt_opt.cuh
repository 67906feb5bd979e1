// The optimiser of pclomp::NormalDistributionsTransform as a resumable state machine, one source for the host and the
// device: computeTransformation (NDT:80-171: Newton step, exit rules), computeStepLengthMT / updateIntervalMT /
// trialValueSelectionMT (NDT:771-931, 647-685, 688-768: More-Thuente line search), computeAngleDerivatives
// (NDT:288-394) and convertTransform (NDT.h:214-231).
//
// An align is a chain of derivative evaluations, each of which needs the sums of the one before it to choose its pose.
// The machine is fed the sums of the pending evaluation (advance) and answers with the next evaluation to run - a pose, its
// f32 transform and angular tables, and a mode - or with "done".  Driven
//   * by the host (ndt.cu, one kernel launch per evaluation): small clouds, profiling, the parity hooks;
//   * by CTA 0 of ndt_align_kernel (ndt_deriv.cuh): the whole align is ONE launch, the grid stays resident, and nothing
//     crosses PCIe between the first evaluation and the result record.
// All state is f64 and every expression keeps the reference's operation order (no FMA contraction: -fmad=false), so both
// drivers walk the same path.
#pragma once
#include "math.cuh"

namespace lgs {
namespace ndtopt {

// ---- trigonometry ------------------------------------------------------------------------------------------------
// The reference takes sinf / cosf of the f32 angles (Eigen::AngleAxisf, NDT.h:222-224) and sin / cos of the f64 angles
// (NDT:293-326).  Here the f32 values are the f64 functions of the widened f32 angle rounded to f32, i.e. the correctly
// rounded sinf / cosf: one definition that the host's libm and the device's libdevice both deliver (their f64 sin / cos
// differ by an ulp now and then, which the rounding to f32 absorbs), so that the two drivers build the same transform.
// A libm whose sinf is not correctly rounded (glibc's is within 0.56 ulp) may differ from this in the last bit.
struct Trig {
  float sf[3], cf[3];   // sin / cos of (float) roll, pitch, yaw, rounded to f32
  double sd[3], cd[3];  // sin / cos of the f64 angles with the small-angle snap of NDT:293-326
};

LGS_HD void trig_of_pose(const double x[6], Trig* t) {
  for (int a = 0; a < 3; a++) {
    const double af = static_cast<double>(static_cast<float>(x[3 + a]));
    t->sf[a] = static_cast<float>(sin(af));
    t->cf[a] = static_cast<float>(cos(af));
    if (fabs(x[3 + a]) < 10e-5) {  // NDT:293-326
      t->cd[a] = 1.0;
      t->sd[a] = 0.0;
    } else {
      t->cd[a] = cos(x[3 + a]);
      t->sd[a] = sin(x[3 + a]);
    }
  }
}

// Eigen::AngleAxis<float>(angle, Unit{X,Y,Z}).toRotationMatrix() from the angle's sine and cosine: note (1-c)*1 + c on
// the axis diagonal
LGS_HD void angle_axis_unit(float s, float c, int axis, float* R) {
  float ax[3] = {0, 0, 0};
  ax[axis] = 1.0f;
  const float sa[3] = {s * ax[0], s * ax[1], s * ax[2]};
  const float ca[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
  float tmp = ca[0] * ax[1];
  R[1] = tmp - sa[2];
  R[3] = tmp + sa[2];
  tmp = ca[0] * ax[2];
  R[2] = tmp + sa[1];
  R[6] = tmp - sa[1];
  tmp = ca[1] * ax[2];
  R[5] = tmp - sa[0];
  R[7] = tmp + sa[0];
  for (int a = 0; a < 3; a++) R[a * 4] = ca[a] * ax[a] + c;
}

LGS_HD void mul3f(const float* a, const float* b, float* c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[i * 3 + j] = (a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j]) + a[i * 3 + 2] * b[6 + j];
}

// NDT.h:214-231: Translation * AngleAxis(X) * AngleAxis(Y) * AngleAxis(Z) in f32, column-major out
LGS_HD void pose_to_matrix(const double x[6], const Trig& t, float* T) {
  float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
  angle_axis_unit(t.sf[0], t.cf[0], 0, Rx);
  angle_axis_unit(t.sf[1], t.cf[1], 1, Ry);
  angle_axis_unit(t.sf[2], t.cf[2], 2, Rz);
  mul3f(Rx, Ry, Rxy);
  mul3f(Rxy, Rz, R);
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[c * 4 + r] = R[r * 3 + c];
  T[12] = static_cast<float>(x[0]);
  T[13] = static_cast<float>(x[1]);
  T[14] = static_cast<float>(x[2]);
}

// computeAngleDerivatives (NDT:329-392): rows a..h of j_ang and a2..f3 of h_ang, f64 and their f32 casts
LGS_HD void angle_tables(const Trig& t, double (*Jd)[3], double (*Hd)[3], float (*Jf)[3], float (*Hf)[3]) {
  const double cx = t.cd[0], cy = t.cd[1], cz = t.cd[2], sx = t.sd[0], sy = t.sd[1], sz = t.sd[2];
  const double J[8][3] = {{(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},
                          {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},
                          {(-sy * cz), sy * sz, cy},
                          {sx * cy * cz, (-sx * cy * sz), sx * sy},
                          {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},
                          {(-cy * sz), (-cy * cz), 0},
                          {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},
                          {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};
  const double H[15][3] = {{(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},
                           {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)},
                           {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},
                           {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},
                           {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},
                           {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},
                           {(-cy * cz), (cy * sz), (sy)},
                           {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},
                           {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},
                           {(sy * sz), (sy * cz), 0},
                           {(-sx * cy * sz), (-sx * cy * cz), 0},
                           {(cx * cy * sz), (cx * cy * cz), 0},
                           {(-cy * cz), (cy * sz), 0},
                           {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},
                           {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 3; c++) {
      Jd[r][c] = J[r][c];
      Jf[r][c] = static_cast<float>(J[r][c]);
    }
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 3; c++) {
      Hd[r][c] = H[r][c];
      Hf[r][c] = static_cast<float>(H[r][c]);
    }
}

// ---- More-Thuente pieces ---------------------------------------------------------------------------------------------
// std::min / std::max as the reference's expressions evaluate them (the argument order decides what a NaN does)
LGS_HD double std_min(double a, double b) { return (b < a) ? b : a; }
LGS_HD double std_max(double a, double b) { return (a < b) ? b : a; }

LGS_HD double dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; i++) s += a[i] * b[i];
  return s;
}

// updateIntervalMT (NDT:647-685)
LGS_HD bool update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    a_u = a_t; f_u = f_t; g_u = g_t;
    return false;
  }
  if (g_t * (a_l - a_t) > 0) {
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  }
  if (g_t * (a_l - a_t) < 0) {
    a_u = a_l; f_u = f_l; g_u = g_l;
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  }
  return true;
}

LGS_HD double mt_cubic(double a0, double f0, double g0, double a1, double f1, double g1) {
  const double z = 3 * (f1 - f0) / (a1 - a0) - g1 - g0;
  const double w = sqrt(z * z - g1 * g0);
  return a0 + (a1 - a0) * (w - g0 - z) / (g1 - g0 + 2 * w);
}

// trialValueSelectionMT (NDT:688-768)
LGS_HD double trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    const double a_c = mt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return fabs(a_c - a_l) < fabs(a_q - a_l) ? a_c : 0.5 * (a_q + a_c);
  }
  if (g_t * g_l < 0) {
    const double a_c = mt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return fabs(a_c - a_t) >= fabs(a_s - a_t) ? a_c : a_s;
  }
  if (fabs(g_t) <= fabs(g_l)) {
    const double a_c = mt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    const double a_next = fabs(a_c - a_t) < fabs(a_s - a_t) ? a_c : a_s;
    const double lim = a_t + 0.66 * (a_u - a_t);
    return a_t > a_l ? std_min(lim, a_next) : std_max(lim, a_next);
  }
  return mt_cubic(a_u, f_u, g_u, a_t, f_t, g_t);
}

// ---- Newton step ------------------------------------------------------------------------------------------------------
// delta = -H^-1 g (NDT:127-129).  The reference goes through Eigen's two-sided JacobiSVD: ~80 dependent plane rotations,
// each a chain of f64 divisions and square roots - 5 us on a CPU core, but 30-40 us on one GPU thread, more than the
// derivative evaluation it sits between.  The solution only enters the align through p -> (f32 transform, f32 tables), so
// the machine solves the well-conditioned case - every regular registration - by Gaussian elimination with partial
// pivoting (equal to the SVD solution to ~cond(H) * 1e-16 relative, orders below the f32 quantisation of the transform)
// and keeps the JacobiSVD, with Eigen's rank threshold, for the ill-conditioned / singular case (pivot ratio below 1e-8,
// non-finite entries), where the minimum-norm semantics of the SVD matter.  Both drivers use this same rule.
LGS_HD bool lu_solve6(const double* A, const double* b, double* x) {
  double M[6][7];
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) M[r][c] = A[r * 6 + c];
    M[r][6] = b[r];
  }
  double pmax = 0, pmin = 0;
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double big = fabs(M[k][k]);
    for (int r = k + 1; r < 6; r++) {
      const double v = fabs(M[r][k]);
      if (v > big) {
        big = v;
        piv = r;
      }
    }
    if (!(big > 0) || big != big || big > 1.7976931348623157e308) return false;
    if (piv != k)
      for (int c = k; c < 7; c++) {
        const double t = M[k][c];
        M[k][c] = M[piv][c];
        M[piv][c] = t;
      }
    if (k == 0) {
      pmax = pmin = big;
    } else {
      pmax = big > pmax ? big : pmax;
      pmin = big < pmin ? big : pmin;
    }
    const double inv = 1.0 / M[k][k];
    for (int r = k + 1; r < 6; r++) {
      const double f = M[r][k] * inv;
      for (int c = k + 1; c < 7; c++) M[r][c] = M[r][c] - f * M[k][c];
    }
  }
  if (!(pmin > 1e-8 * pmax)) return false;
  for (int r = 5; r >= 0; r--) {
    double acc = M[r][6];
    for (int c = r + 1; c < 6; c++) acc = acc - M[r][c] * x[c];
    x[r] = acc / M[r][r];
  }
  for (int r = 0; r < 6; r++)
    if (x[r] != x[r]) return false;
  return true;
}

LGS_HD void newton_solve(const double* H, const double* neg_g, double* delta) {
  if (!lu_solve6(H, neg_g, delta)) m::svd_solve<6>(H, neg_g, delta);
}

// ---- the machine ------------------------------------------------------------------------------------------------------
struct Command {       // the evaluation to run next
  int mode;            // 0: score + g + H (f32 terms), 1: score + g, 2: computeHessian in f64; < 0: the align is done
  float T[16];         // column-major transform applied to the source points
  float j_ang[8][3];   // NDT:339-346, f32
  float h_ang[15][3];  // NDT:373-392, f32
  double j_ang_d[8][3];
  double h_ang_d[15][3];
};

struct Machine {
  // parameters
  double step_size, trans_eps, n_in;
  int max_iter;
  // optimiser state (NDT:103-171)
  double p[6], score, g[6], H[36];
  int nr_iterations, converged, early_exit;
  double trans_probability;
  int evals, trials, hess_recomputes;
  float final_T[16];
  // line-search state (NDT:771-931)
  double dir[6], x_t[6];
  double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, step_min, step_max;
  double phi_t, d_phi_t, psi_t, d_psi_t;
  int interval_converged, open_interval, step_iterations;
  int pending;  // what the evaluation in flight is: 0 initial (NDT:119), 1 first of a line search, 2 trial, 3 computeHessian

  // pose + transform + tables of the evaluation at x
  LGS_HD void pose_command(const double x[6], int mode, Command* c) {
    Trig t;
    trig_of_pose(x, &t);
    pose_to_matrix(x, t, final_T);
    for (int i = 0; i < 16; i++) c->T[i] = final_T[i];
    angle_tables(t, c->j_ang_d, c->h_ang_d, c->j_ang, c->h_ang);
    c->mode = mode;
  }

  // NDT:103-119: the first evaluation, at the pose of the guess (T0 = the guess itself, p0 its translation + Euler angles)
  LGS_HD void begin(const double p0[6], const float T0[16], double step_size_, double trans_eps_, int max_iter_, double n_in_, Command* c) {
    step_size = step_size_;
    trans_eps = trans_eps_;
    max_iter = max_iter_;
    n_in = n_in_;
    for (int i = 0; i < 6; i++) p[i] = p0[i];
    for (int i = 0; i < 16; i++) final_T[i] = T0[i];
    score = 0;
    for (int i = 0; i < 6; i++) g[i] = 0;
    for (int i = 0; i < 36; i++) H[i] = 0;
    nr_iterations = converged = early_exit = 0;
    trans_probability = 0;
    evals = trials = hess_recomputes = 0;
    pending = 0;
    Trig t;
    trig_of_pose(p0, &t);
    for (int i = 0; i < 16; i++) c->T[i] = T0[i];
    angle_tables(t, c->j_ang_d, c->h_ang_d, c->j_ang, c->h_ang);
    c->mode = 0;
    evals = 1;
  }

  LGS_HD void take_sums(const double* s, int mode) {
    // mode 0: score, g[6], upper triangle of H (21); mode 1: score, g[6] (H zeroed, NDT:201); mode 2: upper triangle of H
    if (mode == 2) {
      int k = 0;
      for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) H[i * 6 + j] = H[j * 6 + i] = s[k++];
      return;
    }
    score = s[0];
    for (int i = 0; i < 6; i++) g[i] = s[1 + i];
    if (mode == 0) {
      int k = 7;
      for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) H[i * 6 + j] = H[j * 6 + i] = s[k++];
    } else {
      for (int i = 0; i < 36; i++) H[i] = 0;
    }
  }

  // Feeds the sums of the pending evaluation; returns true and fills *c when another evaluation is needed, false when
  // the align is over (results in nr_iterations / converged / trans_probability / final_T / p).  *c must be the object
  // the previous call (or begin) filled: computeHessian re-uses its transform and tables.
  LGS_HD bool advance(const double* sums, Command* c) {
    const double mu = 1.e-4, nu = 0.9;
    const int max_step_iterations = 10;
    bool newton = false, outer = false;
    double step = 0;
    if (pending == 0) {
      take_sums(sums, 0);
      newton = true;
    } else if (pending == 1 || pending == 2) {
      take_sums(sums, pending == 1 ? 0 : 1);
      phi_t = -score;
      d_phi_t = -dot6(g, dir);
      psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
      d_psi_t = d_phi_t - mu * d_phi_0;
      if (pending == 2) {
        if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
          open_interval = 0;
          f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
          g_l = g_l + mu * d_phi_0;
          f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
          g_u = g_u + mu * d_phi_0;
        }
        if (open_interval)
          interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t) ? 1 : 0;
        else
          interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t) ? 1 : 0;
        step_iterations++;
      }
      // the loop condition of NDT:861
      if (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
        trials++;
        if (open_interval)
          a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
        else
          a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
        a_t = std_max(std_min(a_t, step_max), step_min);
        for (int i = 0; i < 6; i++) x_t[i] = p[i] + dir[i] * a_t;
        pose_command(x_t, 1, c);
        evals++;
        pending = 2;
        return true;
      }
      if (step_iterations) {  // NDT:927-928: the line search iterated, re-evaluate the Hessian in f64 at x_t:
        c->mode = 2;          // *c still holds the transform and the tables of x_t (the contract of advance)
        hess_recomputes++;
        pending = 3;
        return true;
      }
      step = a_t;
      outer = true;
    } else {
      take_sums(sums, 2);
      step = a_t;
      outer = true;
    }
    while (true) {
      if (outer) {  // NDT:144-162
        for (int i = 0; i < 6; i++) p[i] = p[i] + dir[i] * step;
        if (nr_iterations > max_iter || (nr_iterations && (fabs(step) < trans_eps))) converged = 1;
        nr_iterations++;
        if (converged) {
          trans_probability = score / n_in;  // NDT:170
          return false;
        }
        newton = true;
        outer = false;
      }
      if (newton) {  // NDT:127-139
        newton = false;
        double neg_g[6], delta_p[6];
        for (int i = 0; i < 6; i++) neg_g[i] = -g[i];
        newton_solve(H, neg_g, delta_p);
        double delta_p_norm = sqrt(dot6(delta_p, delta_p));
        if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
          trans_probability = score / n_in;
          converged = delta_p_norm == delta_p_norm ? 1 : 0;
          early_exit = 1;
          return false;
        }
        for (int i = 0; i < 6; i++) dir[i] = delta_p[i] / delta_p_norm;
        // computeStepLengthMT (NDT:771-931), up to its first evaluation
        step_max = step_size;
        step_min = trans_eps / 2;
        phi_0 = -score;
        d_phi_0 = -dot6(g, dir);
        if (d_phi_0 >= 0) {
          if (d_phi_0 == 0) {  // NDT:790-791: no step
            step = 0;
            outer = true;
            continue;
          }
          d_phi_0 *= -1;
          for (int i = 0; i < 6; i++) dir[i] *= -1;
        }
        step_iterations = 0;
        a_l = 0;
        a_u = 0;
        f_l = phi_0 - phi_0 - mu * d_phi_0 * a_l;  // auxiliaryFunction_PsiMT (NDT.h:430-436)
        g_l = d_phi_0 - mu * d_phi_0;              // auxiliaryFunction_dPsiMT (NDT.h:438-447)
        f_u = phi_0 - phi_0 - mu * d_phi_0 * a_u;
        g_u = d_phi_0 - mu * d_phi_0;
        interval_converged = (step_max - step_min) < 0 ? 1 : 0;
        open_interval = 1;
        a_t = std_max(std_min(delta_p_norm, step_max), step_min);
        for (int i = 0; i < 6; i++) x_t[i] = p[i] + dir[i] * a_t;
        pose_command(x_t, 0, c);
        evals++;
        pending = 1;
        return true;
      }
    }
  }
};

}  // namespace ndtopt
}  // namespace lgs
