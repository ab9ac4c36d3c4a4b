// bvg_blocks.cu -- batched stand-alone building blocks (post_normal, post_normal_sur, ssvs, bvs, stochvol,
// kalman_dk): placeholder translation unit, filled in below.
#include <string>
#include "../../include/bvargpu.h"
extern "C" {
int bvg_post_normal(int64_t, int, int, int, const double*, const double*, const double*, const double*, const double*, const double*, uint64_t, int, double*) { return -99; }
int bvg_post_normal_sur(int64_t, int, int, int, const double*, const double*, const double*, int, const double*, const double*, const double*, uint64_t, int, double*) { return -99; }
int bvg_ssvs(int64_t, int, const double*, const double*, const double*, const double*, const int32_t*, int, const double*, uint64_t, double*, double*) { return -99; }
int bvg_bvs(int64_t, int, int, int, const double*, const double*, const double*, const double*, const double*, const double*, const int32_t*, int, const double*, const double*, uint64_t, double*) { return -99; }
int bvg_stochvol(int64_t, int, int, const double*, const double*, const double*, const double*, const double*, int, const double*, const double*, uint64_t, double*, int32_t*) { return -99; }
int bvg_kalman_dk(int64_t, int, int, int, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, uint64_t, double*) { return -99; }
}
