"""Load tests/golden/abpoa_golden.npz (made by tests/golden/make_golden.py from the unmodified abPOA)."""
import os

import numpy as np

from oracle.oracle import Dump, PdParams
from smoothxg_b200.synth import PoaBatch

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "abpoa_golden.npz")


def load_cases():
    z = np.load(PATH)
    out = []
    for name in [str(n) for n in z["names"]]:
        batch = PoaBatch(z[f"{name}/bso"], z[f"{name}/sl"], z[f"{name}/so"], z[f"{name}/ba"], z[f"{name}/wt"])
        p = [int(x) for x in z[f"{name}/params"]]
        dumps = [Dump(z[f"{name}/dump{i}"]) for i in range(batch.n_blocks)]
        out.append((name, batch, p, dumps))
    return out


def pd_params(p):
    return PdParams(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], 0.03, p[8], p[9])


def engine_params(p):
    from smoothxg_b200.engine import PoaParams
    return PoaParams(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], 0.03, p[8], p[9])
