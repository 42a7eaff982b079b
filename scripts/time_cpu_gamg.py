"""CPU time of the CG+multigrid solve of the reference problem with the oracle's plain-C V-cycle
(orc_cg_mg, one thread) over the hierarchy the product's host set-up builds.  The baseline beside
the GPU solve times in profiles/r01_gamg_solve.md.  Usage: python scripts/time_cpu_gamg.py N"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostlib  # noqa: E402
import oracle  # noqa: E402
import test_pcgamg as T  # noqa: E402
from oracle import gamg  # noqa: E402

N = int(sys.argv[1])
s = hostlib.System(N)
t0 = time.time()
sv = T.Solver(s.A, {"-pc_gamg_b200_esteig": "gershgorin"})
assert sv.setup_rc == 0
t_setup = time.time() - t0
lv = sv.levels()
rhs = hostlib.vec_array(s.rhs, N ** 3)
exact = hostlib.vec_array(s.exact, N ** 3)
t0 = time.time()
x, its, rn = gamg.cg_mg(lv, rhs)
t_solve = time.time() - t0
ai, aj, aa = s.csr()
t0 = time.time()
xj, its_j, rn_j = oracle.cg_jacobi(ai, aj, aa, rhs)
t_jac = time.time() - t0
print(f"N={N} cores=1 setup_s={t_setup:.2f} (host set-up, {os.cpu_count()} threads) cg_mg: its={its} solve_s={t_solve:.2f} "
      f"err_inf={np.abs(x - exact).max():.3e} | cg_jacobi: its={its_j} solve_s={t_jac:.2f}")
