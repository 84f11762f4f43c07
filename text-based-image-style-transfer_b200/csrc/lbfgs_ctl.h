// L-BFGS controller: the scalar half of torch.optim.LBFGS.step (no line search), written once for
// host and device.  It follows, statement by statement,
//   torch/optim/lbfgs.py:333-537 (torch 2.11; the reference pins 1.13.1, same algorithm), as used by
//   multi_style_transfer/run_style_transfer.py:90,151 (optim.LBFGS([optim_img]) with all defaults).
//
// The vector half lives in lbfgs.cu: one streaming pass over the stored (s, y) pairs yields every dot
// product the two-loop recursion needs ("pass 1"), this controller then runs the recursion in
// coefficient space (the direction is a linear combination of the stored pairs and the gradient),
// and a second streaming pass forms d and applies x <- clamp(x + t d, 0, 1) ("pass 2").
//
// Which dot products the recursion needs (i, j = age order, oldest first):
//   loop 1 (lbfgs.py:432-435)   s_i . q,  q = -g - sum_{j>i} al_j y_j      ->  s_i.g,  s_i.y_j for j > i
//   loop 2 (lbfgs.py:440-442)   y_i . r,  r = H q + sum_{j<i} c_j s_j      ->  y_i.g,  y_i.y_j (all j),  s_j.y_i for j < i
// i.e. the upper triangle R[i][j] = s_i.y_j (j >= i), the symmetric YY[i][j] = y_i.y_j, and Sg, Yg, gg -
// the same quantities as the compact representation of Byrd, Nocedal and Schnabel.  s_i.s_j is never needed.
// Algebraically identical to the reference's sequential recursion; rounding differs (dots reduced in fp64).
//
// Device: one thread block executes nst_lbfgs_control(); the two dependent recurrences run on warp 0 out of
// shared memory, everything else is block-parallel.  Host (unit tests, tests/test_lbfgs_ctl.py drive it against
// torch.optim.LBFGS): a single thread.
#pragma once
#include <stdint.h>
#include <math.h>

#define NST_LBFGS_HISTORY 100
#define NST_LBFGS_SLOTS (NST_LBFGS_HISTORY + 1)  // one spare slot receives the candidate pair
#define NST_LBFGS_NB (2 * NST_LBFGS_SLOTS + 1)   // basis: S slots, Y slots, g
#define NST_LBFGS_G (2 * NST_LBFGS_SLOTS)
#define NST_LBFGS_NSCAL 8
#define NST_LBFGS_NDOT 4  // per stored pair p: S_p.y  S_p.g  Y_p.y  Y_p.g   (y, g = new difference / gradient)

// stop reasons (ctl.stop)
#define NST_RUN 0
#define NST_STOP_OPT_ENTRY 1   // lbfgs.py:369  gradient already below tolerance_grad at step() entry
#define NST_STOP_GTD 2         // lbfgs.py:463  directional derivative above -tolerance_change
#define NST_STOP_OPT 3         // lbfgs.py:518  opt_cond after the update's evaluation
#define NST_STOP_STEP 4        // lbfgs.py:522  max|t d| <= tolerance_change
#define NST_STOP_LOSS 5        // lbfgs.py:525  |loss - prev_loss| < tolerance_change
#define NST_STOP_NONFINITE 6   // loss is NaN/Inf (not in the reference; guards the device loop)
#define NST_STOP_FROZEN 7      // the member sits this step() out (not in the reference: batches of frames, nst_batch_create)

// controller modes
#define NST_CTL_BEGIN 0  // first controller call of a step(): follows the entry evaluation
#define NST_CTL_MID 1    // follows the evaluation that came after an update inside the same step()

struct NstLbfgsCtl {
  // ---- configuration (torch defaults: lr 1, tolerance_grad 1e-7, tolerance_change 1e-9, history 100)
  double lr, tol_grad, tol_change;
  int history_size;
  int frozen;      // != 0: step() entry sets stop = NST_STOP_FROZEN instead of clearing it - a member of a batch whose own loop has
                   // ended while the other members keep stepping (nst_lbfgs_freeze); zero after nst_lbfgs_init
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
  // ---- per-step bookkeeping
  int stop;           // NST_RUN or a stop reason; cleared at step() entry
  int run_pass2;      // 1 if pass 2 has work after this controller call
  int closure_calls;  // iter[0] of run_style_transfer.py:99,143
  int trace_cap;
  float t_apply;      // step applied to x by pass 2 (0 when only d / prev_grad must be refreshed)
  float pad1;
  double gtd, max_td, gmax, gl1, ys, yy;
  // ---- output of the recursion: d = sum_k coef[k] * basis_k  (S slots, Y slots, g)
  float coef[NST_LBFGS_NB + 3];
  // ---- SM clock at the controller's phase boundaries (device only; tuning aid, tools/ctl_phases.py):
  //  0 entry  1 matrices staged  2 pair accepted / rows initialised  3 loop 1 done  4 y.r initialised  5 loop 2 done  6 exit
  long long clk[8];
};

// scalars produced by pass 1 (index into `scal`):  0 s.s  1 s.y  2 y.y  3 s.g  4 y.g  5 g.g  6 max|g|  7 sum|g|

// working arrays of one controller call (shared memory on the device, heap on the host)
struct NstCtlWork {
  double* R;   // [SLOTS][SLOTS]  R[p_i][p_j] = s_i . y_j, valid for age(j) >= age(i)
  double* YY;  // [SLOTS][SLOTS]  symmetric
  double* Sg;  // [SLOTS] by physical slot
  double* Yg;  // [SLOTS]
  double* al;  // [SLOTS]
  double* c;   // [SLOTS] coefficient of s_i in d
  double* yq;  // [SLOTS] y_i . q
  double* ro;  // [SLOTS] 1 / (y_i . s_i)
  double* red; // [4] broadcast scratch
};
// doubles per dot-product matrix: SLOTS^2 rounded up to a multiple of 2 (16-byte granules for bulk copies)
#define NST_CTL_MAT_DOUBLES ((NST_LBFGS_SLOTS * NST_LBFGS_SLOTS + 1) / 2 * 2)
#define NST_CTL_WORK_DOUBLES (2 * NST_CTL_MAT_DOUBLES + 6 * NST_LBFGS_SLOTS + 4)

#if defined(__CUDA_ARCH__)
#if defined(NST_INSTRUMENT)
#define NST_CLK(k) do { if (NST_TID == 0) c->clk[k] = clock64(); } while (0)
#else
#define NST_CLK(k) do { } while (0)
#endif
#define NST_HD __device__
#define NST_TID (static_cast<int>(threadIdx.x))
#define NST_NT (static_cast<int>(blockDim.x))
#define NST_BLOCK_SYNC() __syncthreads()
#define NST_CTL_NL 32
#define NST_WARP_SYNC() __syncwarp()
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
#define NST_CLK(k) ((void)0)
#define NST_HD
#define NST_TID 0
#define NST_NT 1
#define NST_BLOCK_SYNC() ((void)0)
#define NST_CTL_NL 1
#define NST_WARP_SYNC() ((void)0)
static inline double nst_ctl_sum(double v) { return v; }
static inline double nst_ctl_max(double v) { return v; }
#endif

NST_HD inline int nst_ctl_slot(int head, int i) {
  int p = head + i;
  return p >= NST_LBFGS_SLOTS ? p - NST_LBFGS_SLOTS : p;
}

// Rg / YYg: the persistent copies (global memory on the device; may alias w.R / w.YY on the host).
// dots: [SLOTS][NST_LBFGS_NDOT] reduced by pass 1, scal: [NSCAL], td_part: per-block max|t d| of the previous pass 2.
// Must be called by every thread of the block (device) with w.R / w.YY already holding the persistent matrices.
NST_HD inline void nst_lbfgs_control(NstLbfgsCtl* c, NstCtlWork w, double* Rg, double* YYg, const double* dots,
                                     const double* scal, float eval_loss, const float* td_part, int n_td, int mode) {
  const int tid = NST_TID, nt = NST_NT;
  const int TOT = NST_LBFGS_SLOTS;
  if (c->stop != NST_RUN) {
    NST_BLOCK_SYNC();
    if (tid == 0) c->run_pass2 = 0;
    return;
  }
  const double loss = static_cast<double>(eval_loss);
  const double gmax = scal[6];
  if (mode == NST_CTL_MID) {
    if (tid < NST_CTL_NL) {
      double mtd = 0.0;
      for (int i = tid; i < n_td; i += NST_CTL_NL) mtd = fmax(mtd, static_cast<double>(td_part[i]));
      mtd = nst_ctl_max(mtd);
      if (tid == 0) w.red[0] = mtd;
    }
    NST_BLOCK_SYNC();
  }
  int stop = NST_RUN;
  if (!(loss == loss) || isinf(loss)) stop = NST_STOP_NONFINITE;
  if (mode == NST_CTL_BEGIN) {
    // lbfgs.py:364-371
    if (stop == NST_RUN && gmax <= c->tol_grad) stop = NST_STOP_OPT_ENTRY;
  } else if (stop == NST_RUN) {
    // lbfgs.py:511-526 for the iteration whose update was just evaluated
    const double mtd = w.red[0];
    if (gmax <= c->tol_grad) stop = NST_STOP_OPT;
    else if (mtd <= c->tol_change) stop = NST_STOP_STEP;
    else if (fabs(loss - c->prev_loss) < c->tol_change) stop = NST_STOP_LOSS;
  }
  // every thread has read the state it needs for the decision above before thread 0 modifies it
  const int n_iter = c->n_iter + 1;
  int len = c->hist_len, head = c->hist_head;
  double H_diag = c->H_diag;
  const int hist_cap = c->history_size;
  const double lr = c->lr, tol_change = c->tol_change;
  NST_BLOCK_SYNC();
  if (tid == 0) {
    c->func_evals += 1;
    c->loss = loss;
    c->gmax = gmax;
    c->gl1 = scal[7];
    if (mode == NST_CTL_MID) c->max_td = w.red[0];
    if (stop != NST_RUN) {
      c->stop = stop;
      c->run_pass2 = 0;
    }
  }
  if (stop != NST_RUN) return;

  NST_CLK(1);
  // ---- new iteration (lbfgs.py:388-442)
  for (int k = tid; k < NST_LBFGS_NB; k += nt) c->coef[k] = 0.f;
  NST_BLOCK_SYNC();
  if (n_iter == 1) {
    len = 0;
    head = 0;
    H_diag = 1.0;
  } else {
    const double ys = scal[1], yy = scal[2];
    const int pn = nst_ctl_slot(head, len);  // slot pass 1 wrote the candidate pair to
    for (int i = tid; i < len; i += nt) {
      const int p = nst_ctl_slot(head, i);
      w.Sg[p] = dots[NST_LBFGS_NDOT * p + 1];
      w.Yg[p] = dots[NST_LBFGS_NDOT * p + 3];
    }
#if defined(__CUDA_ARCH__)
    // ===================================== device: no dependent chain =====================================================
    // The two loops of the recursion are triangular solves with R (R_ij = s_i . y_j, j >= i in age order, R_ii = y_i . s_i):
    //   loop 1:  R al = -Sg              loop 2:  R^T c = D al - yq0,  D = diag(R),  yq0_i = H (-Yg_i - sum_j al_j YY_ij)
    // (substitute al_k = ro_k (s_k . q) and c_k = al_k - ro_k (y_k . r) into lbfgs.py:432-442).  Run as substitutions they
    // are 2 x 100 dependent fp64 steps of ~140 cycles each (shuffle + DFMA latency): 7 us per loop whether blocked 32 x 32
    // (r01) or run by a single warp (profiles/r02_lbfgs_controller_experiments.md).  On the device the matrix kept in w.R / Rg
    // is therefore M = R^-1, maintained incrementally - appending a pair appends the column -M r / rho and the diagonal
    // 1 / rho (one matrix-vector product), dropping the oldest pair drops the first row and column (the trailing block of
    // the inverse of a triangular matrix is the inverse of the trailing block) - and both solves become matrix-vector
    // products, al = -M Sg and c = M^T (D al - yq0): block parallel, ~1 us each.  fp64 throughout; algebraically identical.
    if (ys > 1e-10) {
      // lbfgs.py:407-421: accept the pair
      for (int i = tid; i < len; i += nt) {
        const int p = nst_ctl_slot(head, i);
        const double Yy = dots[NST_LBFGS_NDOT * p + 2];
        w.YY[p * TOT + pn] = Yy;
        w.YY[pn * TOT + p] = Yy;
        YYg[p * TOT + pn] = Yy;
        YYg[pn * TOT + p] = Yy;
        w.c[p] = dots[NST_LBFGS_NDOT * p + 0];   // r_i = s_i . y_new (scratch: w.c is rewritten below)
      }
      if (tid == 0) {
        w.YY[pn * TOT + pn] = yy;
        YYg[pn * TOT + pn] = yy;
        w.Sg[pn] = scal[3];
        w.Yg[pn] = scal[4];
        c->ro[pn] = 1.0 / ys;
      }
      NST_BLOCK_SYNC();
      const int first = len == hist_cap ? 1 : 0;   // old_dirs.pop(0): the oldest pair leaves
      // row pn of M: the new pair is the newest, its row holds the diagonal only
      for (int j = tid; j < TOT; j += nt) {
        const double v = j == pn ? 1.0 / ys : 0.0;
        w.R[pn * TOT + j] = v;
        Rg[pn * TOT + j] = v;
      }
      // column pn of M for the retained rows: -(M r)_i / rho, four threads per row
      {
        const int row = first + (tid >> 2), part = tid & 3;
        const bool live = row < len;
        const int pi = nst_ctl_slot(head, live ? row : 0);
        double acc = 0.0;
        if (live) {
          for (int j = row + part; j < len; j += 4) {
            const int pj = nst_ctl_slot(head, j);
            acc += w.R[pi * TOT + pj] * w.c[pj];
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (live && part == 0) {
          const double v = -acc / ys;
          w.R[pi * TOT + pn] = v;
          Rg[pi * TOT + pn] = v;
        }
      }
      if (first) head = nst_ctl_slot(head, 1);
      else len += 1;
      H_diag = ys / yy;
    }
    if (tid == 0) {
      c->ys = ys;
      c->yy = yy;
    }
    NST_BLOCK_SYNC();
    for (int i = tid; i < len; i += nt) {
      const int p = nst_ctl_slot(head, i);
      w.ro[p] = c->ro[p];
    }
    NST_BLOCK_SYNC();
    NST_CLK(2);
    // al = -M Sg: row i sums over the columns j >= i, four threads per row
    {
      const int row = tid >> 2, part = tid & 3;
      const bool live = row < len;
      const int pi = nst_ctl_slot(head, live ? row : 0);
      double acc = 0.0;
      if (live) {
        for (int j = row + part; j < len; j += 4) {
          const int pj = nst_ctl_slot(head, j);
          acc += w.R[pi * TOT + pj] * w.Sg[pj];
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (live && part == 0) w.al[pi] = -acc;
    }
    NST_BLOCK_SYNC();
    NST_CLK(3);
    // v_i = D_ii al_i - yq0_i with yq0_i = H (-Yg_i - sum_j al_j YY_ij), four threads per row
    {
      const int row = tid >> 2, part = tid & 3;
      const bool live = row < len;
      const int pi = nst_ctl_slot(head, live ? row : 0);
      double acc = 0.0;
      if (live) {
        for (int j = part; j < len; j += 4) {
          const int pj = nst_ctl_slot(head, j);
          acc += w.al[pj] * w.YY[pi * TOT + pj];
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (live && part == 0) w.yq[pi] = w.al[pi] / w.ro[pi] - H_diag * (-w.Yg[pi] - acc);
    }
    NST_BLOCK_SYNC();
    NST_CLK(4);
    // c = M^T v: row i sums over j <= i
    {
      const int row = tid >> 2, part = tid & 3;
      const bool live = row < len;
      const int pi = nst_ctl_slot(head, live ? row : 0);
      double acc = 0.0;
      if (live) {
        for (int j = part; j <= row; j += 4) {
          const int pj = nst_ctl_slot(head, j);
          acc += w.R[pj * TOT + pi] * w.yq[pj];
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (live && part == 0) w.c[pi] = acc;
    }
    NST_BLOCK_SYNC();
    NST_CLK(5);
#else
    // ===================================== host (unit tests): the recursion as torch writes it ============================
    if (ys > 1e-10) {
      // lbfgs.py:407-421: accept the pair; its dot products with every retained pair enter R / YY
      for (int i = tid; i < len; i += nt) {
        const int p = nst_ctl_slot(head, i);
        const double Sy = dots[NST_LBFGS_NDOT * p + 0], Yy = dots[NST_LBFGS_NDOT * p + 2];
        w.R[p * TOT + pn] = Sy;
        Rg[p * TOT + pn] = Sy;
        w.YY[p * TOT + pn] = Yy;
        w.YY[pn * TOT + p] = Yy;
        YYg[p * TOT + pn] = Yy;
        YYg[pn * TOT + p] = Yy;
      }
      if (tid == 0) {
        w.R[pn * TOT + pn] = ys;
        Rg[pn * TOT + pn] = ys;
        w.YY[pn * TOT + pn] = yy;
        YYg[pn * TOT + pn] = yy;
        w.Sg[pn] = scal[3];
        w.Yg[pn] = scal[4];
        c->ro[pn] = 1.0 / ys;
      }
      if (len == hist_cap) head = nst_ctl_slot(head, 1);  // old_dirs.pop(0)
      else len += 1;
      H_diag = ys / yy;
    }
    if (tid == 0) {
      c->ys = ys;
      c->yy = yy;
    }
    NST_BLOCK_SYNC();
    for (int i = tid; i < len; i += nt) {
      const int p = nst_ctl_slot(head, i);
      w.ro[p] = c->ro[p];
      w.c[p] = -w.Sg[p];  // running value of s_i . q
    }
    NST_BLOCK_SYNC();
    NST_CLK(2);
    // lbfgs.py:432-435: for k newest..oldest: al_k = ro_k (s_k . q); q -= al_k y_k.  Column oriented: once al_k is
    // known every older row i subtracts al_k (s_i . y_k) from its running s_i . q - no reduction on the dependent chain.
    if (tid < NST_CTL_NL) {
      for (int k = len - 1; k >= 0; --k) {
        const int pk = nst_ctl_slot(head, k);
        const double al = w.ro[pk] * w.c[pk];
        for (int i = tid; i < k; i += NST_CTL_NL) {
          const int p = nst_ctl_slot(head, i);
          w.c[p] -= al * w.R[p * TOT + pk];
        }
        if (tid == 0) w.al[pk] = al;
        NST_WARP_SYNC();
      }
    }
    NST_BLOCK_SYNC();
    NST_CLK(3);
    // y_i . r at the start of loop 2: H (y_i . q) = H (-(y_i . g) - sum_j al_j (y_i . y_j))          [block parallel]
    for (int i = tid; i < len; i += nt) {
      const int p = nst_ctl_slot(head, i);
      double acc = 0.0;
      for (int j = 0; j < len; ++j) {
        const int pj = nst_ctl_slot(head, j);
        acc += w.al[pj] * w.YY[p * TOT + pj];
      }
      w.yq[p] = H_diag * (-w.Yg[p] - acc);
    }
    NST_BLOCK_SYNC();
    NST_CLK(4);
    // lbfgs.py:439-442: r = H q; for k oldest..newest: be_k = ro_k (y_k . r); r += (al_k - be_k) s_k.  Once c_k is known
    // every younger row i adds c_k (s_k . y_i) to its running y_i . r.
    if (tid < NST_CTL_NL) {
      for (int k = 0; k < len; ++k) {
        const int pk = nst_ctl_slot(head, k);
        const double ck = w.al[pk] - w.ro[pk] * w.yq[pk];
        for (int i = k + 1 + tid; i < len; i += NST_CTL_NL) {
          const int p = nst_ctl_slot(head, i);
          w.yq[p] += ck * w.R[pk * TOT + p];
        }
        if (tid == 0) w.c[pk] = ck;
        NST_WARP_SYNC();
      }
    }
    NST_BLOCK_SYNC();
    NST_CLK(5);

#endif
    for (int i = tid; i < len; i += nt) {
      const int p = nst_ctl_slot(head, i);
      c->coef[p] = static_cast<float>(w.c[p]);
      c->coef[TOT + p] = static_cast<float>(-H_diag * w.al[p]);
    }
  }
  // lbfgs.py:454-460: step length and directional derivative g.d
  const double t = n_iter == 1 ? fmin(1.0, 1.0 / scal[7]) * lr : lr;
  if (tid < NST_CTL_NL) {
    double acc = 0.0;
    if (n_iter != 1) {
      for (int i = tid; i < len; i += NST_CTL_NL) {
        const int p = nst_ctl_slot(head, i);
        acc += w.c[p] * w.Sg[p] - H_diag * w.al[p] * w.Yg[p];
      }
    }
    const double gtd = nst_ctl_sum(acc) - H_diag * scal[5];
    if (tid == 0) {
      const int gtd_stop = gtd > -tol_change;
      c->coef[NST_LBFGS_G] = static_cast<float>(-H_diag);
      c->n_iter = n_iter;
      c->hist_len = len;
      c->hist_head = head;
      c->H_diag = H_diag;
      c->prev_loss = loss;  // lbfgs.py:448
      c->t = t;
      c->gtd = gtd;
      // pass 2 always refreshes d and prev_flat_grad (lbfgs.py:444-447 run before the gtd test); on a gtd stop it
      // must not move x (lbfgs.py:463-464)
      c->run_pass2 = 1;
      c->t_apply = gtd_stop ? 0.f : static_cast<float>(t);
      if (gtd_stop) c->stop = NST_STOP_GTD;
    }
  }
  NST_BLOCK_SYNC();
  NST_CLK(6);
}
