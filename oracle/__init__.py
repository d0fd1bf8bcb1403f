"""CPU oracle for the 2d_multigrid hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference's algorithm
(`/root/reference/code/6_ntl-mg_new_code/3_combining_laplace_and_wilson/*.h`, abbreviated `S6/`,
and `code/2_scalar_2d_nontelescoping/telescoping_2d_laplace_Mgrid.cpp`, abbreviated `S2/`).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product package (`2d_multigrid_b200/`) never does: it fails loudly if
its CUDA library is missing.

Pinning status (see DESIGN.md section "Oracle"):
  * scalar path (`scalar_s2.py`): pinned by the reference's own golden iteration counts
    (NB/2c_analysis_mass_variation_non-telescoping.ipynb:555-598) and by the S2 binary compiled from
    /root/reference by `oracle/Makefile` into `oracle/_ref/`.
  * Wilson / gauged-Laplace adaptive MG (`mg_oracle.py`): the reference ships no golden vectors and no
    gauge configurations.  It is pinned by (i) the reference's in-run property tests (S6/tests.h),
    (ii) the analytic free Wilson spectrum, and (iii) golden fixtures under tests/golden/ produced by
    compiling the UNMODIFIED S6 sources against a stand-in for the absent Eigen headers
    (`oracle/eigen_shim/`, ours) -- see tests/golden/make_golden.py.
"""
