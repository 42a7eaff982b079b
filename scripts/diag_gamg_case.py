"""One V-cycle application on the device against the C restatement, in its own process (a faulting
kernel poisons the CUDA context).  Usage: python scripts/diag_gamg_case.py N sweeps esteig
Run with CUDA_LAUNCH_BLOCKING=1 to attribute a kernel fault to its launch site."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
import hostlib  # noqa: E402
import test_pcgamg as T  # noqa: E402
from oracle import gamg  # noqa: E402

N, sweeps, esteig = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
import petsc_openacc_b200 as pk  # noqa: E402
pk.init(0)
s = hostlib.System(N)
sv = T.Solver(s.A, {"-pc_gamg_b200_esteig": esteig, "-mg_levels_ksp_max_it": sweeps})
assert sv.setup_rc == 0
lv = sv.levels()
print("levels", [(l["m"], len(l["A"][1])) for l in lv], flush=True)
r = gen.uniform_pm1(N ** 3, 21)
z = T._apply_pc(sv, r)
ref = gamg.mg_apply(lv, r, sweeps=sweeps)
print(f"CASE N={N} sweeps={sweeps} esteig={esteig}: bit-exact={np.array_equal(z, ref)} maxdiff={np.abs(z - ref).max():.3e}", flush=True)
