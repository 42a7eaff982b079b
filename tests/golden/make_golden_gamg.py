"""Regenerates tests/golden/gamg.json from the restatement in oracle/gamg.py + orc_cg_mg (run from
the repo root: `python tests/golden/make_golden_gamg.py`).

The reference records no iteration count anywhere and its PCGAMG is PETSc's, which
is not available here; these numbers pin OUR smoothed-aggregation CG solve of the reference problem
so that a later edit of the set-up or the V-cycle that changes the hierarchy or the count is caught.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import gamg  # noqa: E402


def case(N, esteig):
    p = oracle.poisson7(N)
    lv = gamg.hierarchy(p["ai"], p["aj"], p["aa"], esteig=esteig)
    x, its, rn = gamg.cg_mg(lv, p["rhs"])
    _, its_j, _ = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"])
    return {"N": N, "esteig": esteig, "rows": [int(l["A"].shape[0]) for l in lv], "nnz": [int(l["A"].nnz) for l in lv],
            "emax": [float(l["emax"]) for l in lv[:-1]], "its": int(its), "rnorm": float(rn),
            "error_inf": float(np.abs(x - p["exact"]).max()), "its_cg_jacobi": int(its_j)}


if __name__ == "__main__":
    out = {"made_by": "tests/golden/make_golden_gamg.py",
           "options": "configs/PETSc_SolverOptions_GAMG.info (cg, rtol 1e-14, atol 1e-12; gamg agg nsmooths 1 threshold 0; "
                      "richardson(1)+jacobi levels; jacobi coarse)",
           "cases": [case(8, "gershgorin"), case(16, "gershgorin"), case(16, "cg"), case(24, "gershgorin")]}
    with open(os.path.join(HERE, "gamg.json"), "w") as f:
        json.dump(out, f, indent=1)
    for c in out["cases"]:
        print(c)
