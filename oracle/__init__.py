"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

A literal NumPy/SciPy FP64 restatement of the reference's Gibbs-sampler hot path (same dense matrices, same
operation order, the LAPACK routines Armadillo dispatches to), with every random variate passed in as an array
("injected draws").  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import anything from here; the product (`bvartools_b200`,
`libbvargpu.so`) never does and fails loudly when the CUDA library is missing.

Pinning status: the reference has no tests, golden vectors or fixtures, and neither R nor
Armadillo exists in this image, so the reference cannot be executed here.
  * coefficient draw + inverse-Wishart block (bvaralg.cpp:300-307,511-514; post_normal.cpp:86-93): pinned
    statistically against the only numeric known-answer in the tree, the README summary table
    (tests/golden/readme_table_c1.json), and analytically (posterior mean = OLS, E[Sigma] = SSE/63);
  * SSVS, BVS, KSC/OCSN stochastic volatility, kalman_dk, TVP precision sampler, and the conventions of the
    un-vendored RNG back ends (R nmath rbinom, arma::wishrnd / randg): PARITY UNPINNED -- pinned by source
    reading only.  The injected-variate conventions are documented in oracle/variates.py.
"""
