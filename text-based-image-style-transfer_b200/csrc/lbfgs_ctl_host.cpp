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

// M: [2*SLOTS][2*SLOTS] doubles, v: [2*SLOTS] doubles (caller-owned, zero-initialised)
void nst_ctl_run(NstLbfgsCtl* c, double* M, double* v, const double* dots, const double* scal, float eval_loss,
                 const float* td_part, int n_td, int mode) {
  double cf[NST_LBFGS_NB + 1];
  nst_lbfgs_control(c, M, v, cf, dots, scal, eval_loss, td_part, n_td, mode, 0);
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
