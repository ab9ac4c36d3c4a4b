"""bvartools_b200 -- B200-native Gibbs-sampler hot path behind bvartools' draw_posterior() / bvarpost().

Host-side mirror of the reference interface for that path; the numerics live in csrc/libbvargpu.so (hand-written
sm_100a CUDA kernels behind the C ABI of include/bvargpu.h).  There is no CPU fallback.
"""
from .blocks import (bvs, kalman_dk, loglik_normal, post_normal, post_normal_sur, ssvs, stoch_vol, stochvol_ksc1998,
                     stochvol_ocsn2007)
from .data import e1_readme_sample, load_e1
from .model import add_priors, gen_var, inclusion_prior, minnesota_prior, ssvs_prior
from .posterior import (bvar, bvarpost, draw_posterior, fevd, information_criteria, irf, predict, summary,
                        summary_bvarlist, thin, thin_bvarlist)
from .sampler import ChainSampler, bvaralg, bvartvpalg, pinned_empty, run_chains

__all__ = ["gen_var", "add_priors", "minnesota_prior", "ssvs_prior", "inclusion_prior", "draw_posterior",
           "bvarpost", "bvar", "summary", "bvaralg", "bvartvpalg", "run_chains", "ChainSampler", "pinned_empty",
           "load_e1", "e1_readme_sample", "post_normal", "post_normal_sur", "ssvs", "bvs", "stochvol_ksc1998",
           "stochvol_ocsn2007", "stoch_vol", "kalman_dk", "loglik_normal", "summary_bvarlist", "information_criteria", "thin", "irf", "fevd", "predict", "thin_bvarlist"]
__version__ = "0.1.0"
