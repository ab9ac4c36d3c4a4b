// bvaralg_gpu.cpp -- drop-in for .bvaralg (src/bvaralg.cpp:8, exported as _bvartools_bvaralg, src/RcppExports.cpp:19-27):
// same input list, same return list, the sweeps run in libbvargpu.so.
//
//   .bvaralg_gpu(object)                                  one chain, all defaults: what .bvaralg(object) returns
//   .bvaralg_gpu(object, n_chains = 4096, n_gpus = -1)    a list of 4096 such results, chains sharded over all GPUs
#include "bvg_glue.h"

static int poll_r_interrupt() {            // Rcpp::checkUserInterrupt analogue (bvaralg.cpp:294-296): calling thread only
  try { Rcpp::checkUserInterrupt(); } catch (...) { return 1; }
  return 0;
}

// [[Rcpp::export(.bvaralg_gpu)]]
Rcpp::List bvaralg_gpu(Rcpp::List object, int n_chains = 1, int n_gpus = 1, int device = -1, double seed = -1) {
  bvg_set_interrupt_poll(poll_r_interrupt);
  Rcpp::List res = bvgglue::run(object, false, n_chains, n_gpus, device, seed);
  bvg_set_interrupt_poll(NULL);
  return res;
}
