"""Generates tests/golden/matparams_golden.npz: a seeded scene (tests/problems.py::matparams_scene) -> the smoothed
eps and mu arrays of the ORACLE restatement (oracle/matparams.py; the reference cannot run here and holds no fixture
for calc_matparams!, see the oracle's header).  Freezes the oracle's behaviour so later edits to the oracle or to
csrc/matparams.cu cannot drift silently.  Run from the repo root:  python tests/golden/make_golden_matparams.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from problems import matparams_scene, MATPARAMS_CASES  # noqa: E402
from oracle import matparams as omp  # noqa: E402
from oracle.grid import Grid  # noqa: E402

out = {}
for i in (1, 2):                                   # non-uniform grids, full-tensor materials: eps (mixed boundft) and mu
    N, isbloch, boundft, ft, uniform, nshape, aniso = MATPARAMS_CASES[i]
    lp, o_sh, _, pinds, params = matparams_scene(N, isbloch, uniform, nshape, aniso)
    out[f"case{i}"] = omp.calc_matparams(Grid(lp, isbloch), boundft, ft, o_sh, pinds, params)
np.savez_compressed(os.path.join(HERE, "matparams_golden.npz"), **out)
print("wrote", os.path.join(HERE, "matparams_golden.npz"))
