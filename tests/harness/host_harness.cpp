// host_harness.cpp — TEST-ONLY host build of the engine's host/device-shared logic (d2d_pair.h,
// optimizer.h) so the CPU suite can check it against the oracle without a GPU.  Never shipped,
// never loaded by the product.
#include "../../ndt_feature_graph_b200/csrc/optimizer.h"

#include <cstring>

using namespace ndtb;

extern "C" {

int hh_pair(const double *mu, const double *C6, const double *m, const double *S6, double lfd1, double lfd2,
            int hess, double *acc28, double *g6) {
  bool ok = hess ? pair_contrib<true>(mu[0], mu[1], mu[2], C6, m, S6, lfd1, lfd2, acc28, g6)
                 : pair_contrib<false>(mu[0], mu[1], mu[2], C6, m, S6, lfd1, lfd2, acc28, g6);
  return ok ? 1 : 0;
}

typedef int (*hh_eval_cb)(const double *T16, int want_hess, double *sums28);

struct hh_result {
  double T[16];
  double score, score_best;
  int converged, iterations, n_hess, n_grad, exit_code, nonfinite, n_exec;
};

int hh_match(const double *T0, int itr_max, int step_control, int regularize, double delta_score, int fusion,
             int soft, int tik, const double *Tcov36, hh_eval_cb cb, hh_result *out, int planar) {
  OptParams prm;
  std::memset(&prm, 0, sizeof prm);
  prm.itr_max = itr_max, prm.step_control = step_control, prm.regularize = regularize, prm.planar = planar;
  prm.delta_score = delta_score;
  prm.fusion = fusion, prm.soft = fusion && soft, prm.tik = fusion && tik;
  if (fusion && !inv6(Tcov36, prm.Q)) return -2;
  OptState s;
  opt_begin(s, prm, T0);
  int guard = 0;
  while (s.phase != PH_DONE) {
    double T16[16], sums[28];
    pose_to_cm(s.Peval, T16);
    std::memset(sums, 0, sizeof sums);
    if (cb(T16, s.want_hess, sums) != 0) return -1;
    opt_advance(s, prm, sums);
    if (++guard > 100000) return -3;
  }
  pose_to_cm(s.T, out->T);
  out->score = s.score_here, out->score_best = s.score_best;
  out->converged = s.ret, out->iterations = s.itr, out->n_hess = s.n_hess, out->n_grad = s.n_grad;
  out->exit_code = s.exit_code, out->nonfinite = s.nonfinite, out->n_exec = s.n_exec;
  return 0;
}

int hh_cstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp, double fp,
             double dp, int *brackt, double stmin, double stmax) {
  return mt_cstep(*stx, *fx, *dx, *sty, *fy, *dy, *stp, fp, dp, *brackt, stmin, stmax);
}

void hh_eig_sym(int n, const double *A, double *evals, double *V) { eig_sym(n, A, evals, V); }
void hh_ldlt6(const double *A, const double *b, double *x) { ldlt_solve6(A, b, x); }
int hh_inv6(const double *A, double *Ainv) { return inv6(A, Ainv) ? 1 : 0; }
}
