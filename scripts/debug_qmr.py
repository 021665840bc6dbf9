import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from problems import Problem, rel
p = Problem((8, 7, 6), (False, False, False), npml=2)
A_ref, _ = p.oracle_csc()
b = A_ref.matvec(p.random_x(3))
A = p.operator(device=0)
for method in ("bicgstab", "qmr"):
    for maxit in (1, 2, 5, 400):
        x, info = A.solve(torch.from_numpy(b).cuda(), method=method, rtol=1e-10, maxit=maxit, check_every=1, history=True)
        xs = x.cpu().numpy()
        print(method, maxit, {k: v for k, v in info.items() if k != "history"}, "true", rel(A_ref.matvec(xs), b),
              "hist", info["history"][:4], np.abs(xs).max())
