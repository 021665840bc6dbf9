"""Generates tests/golden/apply_golden.npz: seeded inputs -> oracle outputs (y = A x, Julia CSC pattern).

The reference cannot run here (no Julia; un-vendored MaxwellBase), so these vectors are produced by the
ORACLE restatement (oracle/operators.py), which is itself pinned by tests/test_oracle_properties.py.
They freeze the oracle's behaviour so later edits to the oracle or the kernels cannot drift silently.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from problems import Problem  # noqa: E402

CASES = {
    "bloch_full": dict(N=(9, 7, 6), isbloch=(True, True, True), full_eps=True, with_mu=True),
    "sym_diag": dict(N=(8, 5, 7), isbloch=(False, False, False), full_eps=False, with_mu=False),
}

out = {}
for tag, kw in CASES.items():
    p = Problem(**kw)
    A, _ = p.oracle_csc()
    x = p.random_x()
    cp, rv = A.julia_pattern()
    out[f"{tag}_N"] = np.array(kw["N"])
    out[f"{tag}_isbloch"] = np.array(kw["isbloch"])
    out[f"{tag}_full"] = np.array(kw["full_eps"])
    out[f"{tag}_mu"] = np.array(kw["with_mu"])
    out[f"{tag}_x"] = x
    out[f"{tag}_y"] = A.matvec(x)
    out[f"{tag}_colptr"] = cp
    out[f"{tag}_rowval"] = rv
np.savez_compressed(os.path.join(HERE, "apply_golden.npz"), **out)
print("wrote", os.path.join(HERE, "apply_golden.npz"))
