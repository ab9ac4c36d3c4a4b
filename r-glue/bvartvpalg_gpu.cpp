// bvartvpalg_gpu.cpp -- drop-in for .bvartvpalg (src/bvartvpalg.cpp:10, exported as _bvartools_bvartvpalg,
// src/RcppExports.cpp:30-38): same input list, same return list (TVP coefficient paths time-major, state variances under
// $sigma of every block, Sigma_t per period with SV or the covariance block), the sweeps run in libbvargpu.so.
#include "bvg_glue.h"

static int poll_r_interrupt_tvp() {
  try { Rcpp::checkUserInterrupt(); } catch (...) { return 1; }
  return 0;
}

// [[Rcpp::export(.bvartvpalg_gpu)]]
Rcpp::List bvartvpalg_gpu(Rcpp::List object, int n_chains = 1, int n_gpus = 1, int device = -1, double seed = -1) {
  bvg_set_interrupt_poll(poll_r_interrupt_tvp);
  Rcpp::List res = bvgglue::run(object, true, n_chains, n_gpus, device, seed);
  bvg_set_interrupt_poll(NULL);
  return res;
}
