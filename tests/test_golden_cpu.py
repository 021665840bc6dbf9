"""The committed golden vectors still equal what the oracle (CSC and matrix-free) computes today, and the
host-only pattern export of the product reproduces the committed Julia CSC pattern."""
import os

import numpy as np

from problems import Problem, rel


def test_oracle_reproduces_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "apply_golden.npz"))
    for tag in ("bloch_full", "sym_diag"):
        p = Problem(N=tuple(g[f"{tag}_N"]), isbloch=tuple(bool(b) for b in g[f"{tag}_isbloch"]),
                    full_eps=bool(g[f"{tag}_full"]), with_mu=bool(g[f"{tag}_mu"]))
        x = p.random_x()
        assert np.array_equal(x, g[f"{tag}_x"])
        A, _ = p.oracle_csc()
        assert rel(A.matvec(x), g[f"{tag}_y"]) < 1e-14
        assert rel(p.oracle_matfree()(x), g[f"{tag}_y"]) < 1e-13
        cp, rv = A.julia_pattern()
        assert np.array_equal(cp, g[f"{tag}_colptr"]) and np.array_equal(rv, g[f"{tag}_rowval"])
        cp2, rv2, _ = p.operator(device=-2).export_pattern(values=False)
        assert np.array_equal(cp2, g[f"{tag}_colptr"]) and np.array_equal(rv2, g[f"{tag}_rowval"])


def test_oracle_reproduces_matparams_golden():
    """material pipeline (N4): the committed arrays (tests/golden/make_golden_matparams.py) equal today's oracle"""
    from oracle import matparams as omp
    from oracle.grid import Grid
    from problems import matparams_scene, MATPARAMS_CASES
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "matparams_golden.npz"))
    for i in (1, 2):
        N, isbloch, boundft, ft, uniform, nshape, aniso = MATPARAMS_CASES[i]
        lp, o_sh, _, pinds, params = matparams_scene(N, isbloch, uniform, nshape, aniso)
        assert rel(omp.calc_matparams(Grid(lp, isbloch), boundft, ft, o_sh, pinds, params), g[f"case{i}"]) < 1e-14
