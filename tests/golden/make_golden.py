"""Regenerates tests/golden/*.json from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).

The reference ships no golden vectors; these pin OUR restatement so that any
later edit of the oracle, the product generator or the kernels that changes a single bit of the
50^3 case (BASELINE configs[0]) is caught.  Checksums are sha256 of the raw little-endian arrays.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import gen  # noqa: E402
import oracle  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seq_case(N):
    p = oracle.poisson7(N)
    x = gen.uniform_pm1(N ** 3, seed=0xB200)
    out = {"N": N, "rows": N ** 3, "nnz": int(p["ai"][-1]), "scale": p["scale"].hex(),
           "ai": sha(p["ai"]), "aj": sha(p["aj"]), "aa": sha(p["aa"]), "rhs": sha(p["rhs"]),
           "exact": sha(p["exact"]),
           "y_rand": sha(oracle.matmult(p["ai"], p["aj"], p["aa"], x)),
           "y_rand_fma": sha(oracle.matmult(p["ai"], p["aj"], p["aa"], x, fma=True)),
           "y_exact": sha(oracle.matmult(p["ai"], p["aj"], p["aa"], p["exact"])),
           "yt_rand": sha(oracle.matmulttranspose(p["ai"], p["aj"], p["aa"], x, N ** 3)),
           "offdiag_value": float(p["aa"][p["ai"][1] + 1]).hex()}
    return out


def mpi_case(N, size):
    ranks = []
    base = oracle.dmda_bases(N, N, N, size)
    for r in range(size):
        p = oracle.poisson7(N, size=size, rank=r)
        (Ai, Aj, Aa), (Bi, Bj, Ba) = oracle.mpiaij_split(p["ai"], p["aj"], p["aa"], p["rstart"], p["rend"])
        Bjc, garray = oracle.mpiaij_setup_multiply(Bj)
        off = oracle.scatter_recv_offsets(base, garray)
        ranks.append({"rstart": p["rstart"], "rend": p["rend"], "A_nnz": len(Aj), "B_nnz": len(Bj),
                      "nghost": len(garray), "Ai": sha(Ai), "Aj": sha(Aj), "Aa": sha(Aa),
                      "Bi": sha(Bi), "Bj": sha(Bjc), "Ba": sha(Ba), "garray": sha(garray),
                      "recv_off": [int(v) for v in off]})
    return {"N": N, "size": size, "grid": list(oracle.dmda_decide(N, N, N, size)),
            "bases": [int(b) for b in base], "ranks": ranks}


def synthetic_case():
    """The synthetic workloads of SURVEY 8(d): 27-point box stencil and the power-law matrix."""
    out = {}
    ai, aj, aa = gen.stencil27(9)
    out["stencil27_9"] = {"nnz": len(aj), "ai": sha(ai), "aj": sha(aj), "aa": sha(aa)}
    ai, aj, aa = gen.powerlaw(20000, lmax=3000)
    x = gen.uniform_pm1(20000, seed=0xB200)
    out["powerlaw_20000_3000"] = {"nnz": len(aj), "rmax": int(np.diff(ai).max()), "ai": sha(ai), "aj": sha(aj), "aa": sha(aa),
                                  "y": sha(oracle.matmult(ai, aj, aa, x)), "yt": sha(oracle.matmulttranspose(ai, aj, aa, x, 20000))}
    out["x_b200_16"] = sha(gen.uniform_pm1(16, seed=0xB200))
    return out


def main():
    gold = {"seq": [seq_case(N) for N in (8, 50)], "synthetic": synthetic_case(),
            "mpi": [mpi_case(N, s) for N, s in ((12, 2), (12, 4), (12, 8), (13, 8), (10, 3))],
            "decide_300": {str(s): list(oracle.dmda_decide(300, 300, 300, s)) for s in (1, 2, 4, 8)}}
    with open(os.path.join(HERE, "poisson7.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "poisson7.json"))


if __name__ == "__main__":
    main()
