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
