"""diagnostic: ocean model, Love group velocity through libsurf.forward (vp = 1.732 vs = 0 in the water:
the reference's start value is c = 0) — thread-mapped vs team-mapped root search vs oracle"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle
from rfsurfhmc_b200._lib import Context
thk = np.array([3.0, 2.0, 5.0, 12.0, 0.0]); vs = np.array([0.0, 2.2, 3.3, 3.9, 4.6])
vp = np.array([1.5, 4.2, 5.9, 6.8, 8.1]); rho = np.array([1.03, 2.3, 2.7, 2.9, 3.3])
T = np.array([6., 10., 15., 25., 40.])
O = Oracle(); ctx = Context(0)
for wt in ("Rc", "Rg", "Lc", "Lg"):
    print(wt, "oracle", O.surf_forward(thk, vp, vs, rho, T, wt))
    for ts in ((0, 1), (4, 4), (8, 1), (32, 4)):
        ctx.set_roots_team(*ts)
        print("   ", ts, ctx.surf_forward(thk, vp, vs, rho, T, wt))
