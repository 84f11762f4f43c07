// L-BFGS controller: the scalar half of torch.optim.LBFGS.step (no line search), written once for
// host and device.  It follows, statement by statement,
//   torch/optim/lbfgs.py:333-537 (torch 2.11; the reference pins 1.13.1, same algorithm), as used by
//   multi_style_transfer/run_style_transfer.py:90,151 (optim.LBFGS([optim_img]) with all defaults).
//
// The vector half lives in lbfgs.cu: one streaming pass over the stored (s, y) pairs yields every dot
// product the two-loop recursion needs ("pass 1"), this controller then runs the recursion in
// coefficient space (the direction is a linear combination of the stored pairs and the gradient),
// and a second streaming pass forms d, applies x <- clamp(x + t d, 0, 1) ("pass 2").
// Algebraically identical to the reference's sequential recursion; rounding differs (dots are reduced
// in fp64 here).
//
// On the device one warp executes nst_lbfgs_control(); on the host (unit tests, tests/test_lbfgs_ctl.py
// drive it against torch.optim.LBFGS) a single thread does.
#pragma once
#include <stdint.h>
#include <math.h>

#define NST_LBFGS_HISTORY 100
#define NST_LBFGS_SLOTS (NST_LBFGS_HISTORY + 1)  // one spare slot receives the candidate pair
#define NST_LBFGS_NB (2 * NST_LBFGS_SLOTS + 1)   // basis: S slots, Y slots, g
#define NST_LBFGS_G (2 * NST_LBFGS_SLOTS)
#define NST_LBFGS_NSCAL 8

// stop reasons (ctl.stop)
#define NST_RUN 0
#define NST_STOP_OPT_ENTRY 1   // lbfgs.py:369  gradient already below tolerance_grad at step() entry
#define NST_STOP_GTD 2         // lbfgs.py:463  directional derivative above -tolerance_change
#define NST_STOP_OPT 3         // lbfgs.py:518  opt_cond after the update's evaluation
#define NST_STOP_STEP 4        // lbfgs.py:522  max|t d| <= tolerance_change
#define NST_STOP_LOSS 5        // lbfgs.py:525  |loss - prev_loss| < tolerance_change
#define NST_STOP_NONFINITE 6   // loss is NaN/Inf (not in the reference; guards the device loop)

// controller modes
#define NST_CTL_BEGIN 0  // first controller call of a step(): follows the entry evaluation
#define NST_CTL_MID 1    // follows the evaluation that came after an update inside the same step()

#ifdef __CUDACC__
#define NST_HD __host__ __device__
#else
#define NST_HD
#endif

struct NstLbfgsCtl {
  // ---- configuration (torch defaults: lr 1, tolerance_grad 1e-7, tolerance_change 1e-9, history 100)
  double lr, tol_grad, tol_change;
  int history_size;
  int pad0;
  // ---- persistent optimizer state (torch `state` dict)
  int n_iter;      // state['n_iter']
  int func_evals;  // state['func_evals']
  int hist_len;    // len(old_dirs)
  int hist_head;   // physical slot of the oldest pair; pairs occupy (head + i) % SLOTS, i < len
  double t;        // state['t']
  double H_diag;   // state['H_diag']
  double prev_loss;
  double loss;
  double ro[NST_LBFGS_SLOTS];  // by physical slot
  double al[NST_LBFGS_SLOTS];
  // ---- per-step bookkeeping
  int stop;           // NST_RUN or a stop reason; cleared by the host/`step_begin` at step() entry
  int run_pass2;      // 1 if pass 2 has work after this controller call
  int closure_calls;  // iter[0] of run_style_transfer.py:99,143
  int trace_cap;
  float t_apply;      // step applied to x by pass 2 (0 when only d / prev_grad must be refreshed)
  float pad1;
  double gtd, max_td, gmax, gl1, ys, yy;
  // ---- output of the recursion: d = sum_k coef[k] * basis_k
  float coef[NST_LBFGS_NB + 3];
};

// scalars produced by pass 1 (index into `scal`)
//  0 s.s  1 s.y  2 y.y  3 s.g  4 y.g  5 g.g  6 max|g|  7 sum|g|
// per used physical slot p, `dots` holds 6 numbers: S_p.s S_p.y S_p.g Y_p.s Y_p.y Y_p.g

#if defined(__CUDA_ARCH__)
#define NST_CTL_LANE (static_cast<int>(threadIdx.x) & 31)
#define NST_CTL_NL 32
#define NST_CTL_SYNC() __syncwarp()
__device__ __forceinline__ double nst_ctl_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double nst_ctl_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#else
#define NST_CTL_LANE 0
#define NST_CTL_NL 1
#define NST_CTL_SYNC() ((void)0)
static inline double nst_ctl_sum(double v) { return v; }
static inline double nst_ctl_max(double v) { return v; }
#endif

// M: [2*SLOTS][2*SLOTS] dot products between stored vectors (row-major, physical slot indexing:
//    S_p -> p, Y_p -> SLOTS + p).  v: [2*SLOTS] dots with the current gradient.
// cf: [NB] fp64 scratch for the coefficients (shared memory on the device).
// eval_loss: the closure value; td_part/n_td: per-block max|t d| partials of the previous pass 2.
NST_HD inline void nst_lbfgs_control(NstLbfgsCtl* c, double* M, double* v, double* cf, const double* dots,
                                     const double* scal, float eval_loss, const float* td_part, int n_td, int mode,
                                     int is_last_iter) {
  const int lane = NST_CTL_LANE;
  const int TOT = NST_LBFGS_SLOTS;
  const int LD = 2 * NST_LBFGS_SLOTS;
  if (c->stop != NST_RUN) {
    if (lane == 0) c->run_pass2 = 0;
    NST_CTL_SYNC();
    return;
  }
  const double loss = static_cast<double>(eval_loss);
  const double gmax = scal[6];
  int stop = NST_RUN;
  if (!(loss == loss) || isinf(loss)) stop = NST_STOP_NONFINITE;
  if (mode == NST_CTL_BEGIN) {
    // lbfgs.py:364-371
    if (stop == NST_RUN && gmax <= c->tol_grad) stop = NST_STOP_OPT_ENTRY;
  } else {
    // lbfgs.py:511-526 for the iteration whose update was just evaluated
    double mtd = 0.0;
    for (int i = lane; i < n_td; i += NST_CTL_NL) mtd = fmax(mtd, static_cast<double>(td_part[i]));
    mtd = nst_ctl_max(mtd);
    if (lane == 0) c->max_td = mtd;
    if (stop == NST_RUN) {
      if (gmax <= c->tol_grad) stop = NST_STOP_OPT;
      else if (mtd <= c->tol_change) stop = NST_STOP_STEP;
      else if (fabs(loss - c->prev_loss) < c->tol_change) stop = NST_STOP_LOSS;
    }
  }
  NST_CTL_SYNC();
  if (lane == 0) {
    c->func_evals += 1;
    c->loss = loss;
    c->gmax = gmax;
    c->gl1 = scal[7];
  }
  if (stop != NST_RUN) {
    if (lane == 0) {
      c->stop = stop;
      c->run_pass2 = 0;
    }
    NST_CTL_SYNC();
    return;
  }

  // ---- new iteration (lbfgs.py:388-442)
  const int n_iter = c->n_iter + 1;
  int len = c->hist_len, head = c->hist_head;
  double H_diag = c->H_diag;
  if (n_iter == 1) {
    for (int k = lane; k < NST_LBFGS_NB; k += NST_CTL_NL) cf[k] = 0.0;
    NST_CTL_SYNC();
    if (lane == 0) cf[NST_LBFGS_G] = -1.0;
    len = 0;
    head = 0;
    H_diag = 1.0;
  } else {
    const double ys = scal[1], yy = scal[2];
    const int pn = (head + len) % TOT;  // slot pass 1 wrote the candidate pair to
    // dots of the stored pairs with the gradient
    for (int i = lane; i < len; i += NST_CTL_NL) {
      const int p = (head + i) % TOT;
      v[p] = dots[6 * p + 2];
      v[TOT + p] = dots[6 * p + 5];
    }
    if (ys > 1e-10) {
      // lbfgs.py:407-421: accept the pair; its dots with every retained pair enter M
      for (int i = lane; i < len; i += NST_CTL_NL) {
        const int p = (head + i) % TOT;
        const double Ss = dots[6 * p + 0], Sy = dots[6 * p + 1], Ys = dots[6 * p + 3], Yy = dots[6 * p + 4];
        M[p * LD + pn] = Ss;
        M[pn * LD + p] = Ss;
        M[p * LD + TOT + pn] = Sy;
        M[(TOT + pn) * LD + p] = Sy;
        M[(TOT + p) * LD + pn] = Ys;
        M[pn * LD + TOT + p] = Ys;
        M[(TOT + p) * LD + TOT + pn] = Yy;
        M[(TOT + pn) * LD + TOT + p] = Yy;
      }
      if (lane == 0) {
        M[pn * LD + pn] = scal[0];
        M[pn * LD + TOT + pn] = ys;
        M[(TOT + pn) * LD + pn] = ys;
        M[(TOT + pn) * LD + TOT + pn] = yy;
        v[pn] = scal[3];
        v[TOT + pn] = scal[4];
        c->ro[pn] = 1.0 / ys;
      }
      if (len == c->history_size) head = (head + 1) % TOT;
      else len += 1;
      H_diag = ys / yy;
    }
    if (lane == 0) {
      c->ys = ys;
      c->yy = yy;
    }
    NST_CTL_SYNC();
    // lbfgs.py:432-435: q = -g; for i newest..oldest: al_i = ro_i (s_i . q); q -= al_i y_i
    for (int k = lane; k < NST_LBFGS_NB; k += NST_CTL_NL) cf[k] = 0.0;
    NST_CTL_SYNC();
    if (lane == 0) cf[NST_LBFGS_G] = -1.0;
    NST_CTL_SYNC();
    for (int i = len - 1; i >= 0; --i) {
      const int p = (head + i) % TOT;
      double acc = 0.0;
      for (int k = lane; k < LD; k += NST_CTL_NL) acc += cf[k] * M[p * LD + k];
      acc = nst_ctl_sum(acc) + cf[NST_LBFGS_G] * v[p];
      const double al = c->ro[p] * acc;
      NST_CTL_SYNC();
      if (lane == 0) {
        c->al[p] = al;
        cf[TOT + p] -= al;
      }
      NST_CTL_SYNC();
    }
    // lbfgs.py:439: r = q * H_diag
    for (int k = lane; k < NST_LBFGS_NB; k += NST_CTL_NL) cf[k] *= H_diag;
    NST_CTL_SYNC();
    // lbfgs.py:440-442: for i oldest..newest: be_i = ro_i (y_i . r); r += (al_i - be_i) s_i
    for (int i = 0; i < len; ++i) {
      const int p = (head + i) % TOT;
      double acc = 0.0;
      for (int k = lane; k < LD; k += NST_CTL_NL) acc += cf[k] * M[(TOT + p) * LD + k];
      acc = nst_ctl_sum(acc) + cf[NST_LBFGS_G] * v[TOT + p];
      const double be = c->ro[p] * acc;
      NST_CTL_SYNC();
      if (lane == 0) cf[p] += c->al[p] - be;
      NST_CTL_SYNC();
    }
  }
  NST_CTL_SYNC();
  // lbfgs.py:454-460: step length and directional derivative
  const double t = n_iter == 1 ? fmin(1.0, 1.0 / scal[7]) * c->lr : c->lr;
  double gtd = 0.0;
  for (int k = lane; k < LD; k += NST_CTL_NL) gtd += cf[k] * v[k];
  gtd = nst_ctl_sum(gtd) + cf[NST_LBFGS_G] * scal[5];
  const int gtd_stop = gtd > -c->tol_change;
  for (int k = lane; k < NST_LBFGS_NB; k += NST_CTL_NL) c->coef[k] = static_cast<float>(cf[k]);
  if (lane == 0) {
    c->n_iter = n_iter;
    c->hist_len = len;
    c->hist_head = head;
    c->H_diag = H_diag;
    c->prev_loss = loss;  // lbfgs.py:448
    c->t = t;
    c->gtd = gtd;
    // pass 2 always refreshes d and prev_flat_grad (lbfgs.py:444-447 run before the gtd test);
    // on a gtd stop it must not move x (lbfgs.py:463-464)
    c->run_pass2 = 1;
    c->t_apply = gtd_stop ? 0.f : static_cast<float>(t);
    if (gtd_stop) c->stop = NST_STOP_GTD;
  }
  (void)is_last_iter;
  NST_CTL_SYNC();
}
