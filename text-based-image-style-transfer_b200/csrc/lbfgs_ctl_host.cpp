// Host build of the L-BFGS controller (lbfgs_ctl.h) for CPU unit tests: tests/test_lbfgs_ctl.py drives it
// with numpy standing in for the two streaming passes and compares against torch.optim.LBFGS.
#include "lbfgs_ctl.h"

#include <stdlib.h>
#include <string.h>

extern "C" {

int nst_ctl_sizeof(void) { return static_cast<int>(sizeof(NstLbfgsCtl)); }
int nst_ctl_slots(void) { return NST_LBFGS_SLOTS; }

NstLbfgsCtl* nst_ctl_new(void) {
  NstLbfgsCtl* c = static_cast<NstLbfgsCtl*>(calloc(1, sizeof(NstLbfgsCtl)));
  c->lr = 1.0;
  c->tol_grad = 1e-7;
  c->tol_change = 1e-9;
  c->history_size = NST_LBFGS_HISTORY;
  c->H_diag = 1.0;
  return c;
}
void nst_ctl_free(NstLbfgsCtl* c) { free(c); }

int nst_ctl_ndot(void) { return NST_LBFGS_NDOT; }
int nst_ctl_work_doubles(void) { return NST_CTL_WORK_DOUBLES; }

// work: NST_CTL_WORK_DOUBLES doubles, zero-initialised and kept by the caller between calls (its R / YY parts are the
// persistent dot-product matrices)
void nst_ctl_run(NstLbfgsCtl* c, double* work, const double* dots, const double* scal, float eval_loss,
                 const float* td_part, int n_td, int mode) {
  NstCtlWork w;
  w.R = work;
  w.YY = w.R + NST_CTL_MAT_DOUBLES;
  w.Sg = w.YY + NST_CTL_MAT_DOUBLES;
  w.Yg = w.Sg + NST_LBFGS_SLOTS;
  w.al = w.Yg + NST_LBFGS_SLOTS;
  w.c = w.al + NST_LBFGS_SLOTS;
  w.yq = w.c + NST_LBFGS_SLOTS;
  w.ro = w.yq + NST_LBFGS_SLOTS;
  w.red = w.ro + NST_LBFGS_SLOTS;
  nst_lbfgs_control(c, w, w.R, w.YY, dots, scal, eval_loss, td_part, n_td, mode);
}

void nst_ctl_begin_step(NstLbfgsCtl* c) {
  c->stop = NST_RUN;
  c->run_pass2 = 0;
}

// field accessors (keeps the Python side independent of the struct layout)
int nst_ctl_get_int(const NstLbfgsCtl* c, int which) {
  switch (which) {
    case 0: return c->n_iter;
    case 1: return c->func_evals;
    case 2: return c->hist_len;
    case 3: return c->hist_head;
    case 4: return c->stop;
    case 5: return c->run_pass2;
    default: return -1;
  }
}
double nst_ctl_get_double(const NstLbfgsCtl* c, int which) {
  switch (which) {
    case 0: return c->t;
    case 1: return c->H_diag;
    case 2: return c->prev_loss;
    case 3: return c->loss;
    case 4: return c->gtd;
    case 5: return static_cast<double>(c->t_apply);
    default: return 0.0;
  }
}
const float* nst_ctl_coef(const NstLbfgsCtl* c) { return c->coef; }
}
