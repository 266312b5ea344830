"""Load tests/golden/abpoa_golden.npz (made by tests/golden/make_golden.py from the unmodified abPOA)."""
import os

import numpy as np

from oracle.oracle import Dump, PdParams
from smoothxg_b200.synth import PoaBatch

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "abpoa_golden.npz")


REAL_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "drb1_golden.npz")


def load_real_cases():
    """tests/golden/drb1_golden.npz (tests/golden/make_real_golden.py): the 17 real blocks smoothxg forms from its own test
    data (DRB1-3123, -l 1100), exactly as handed to abPOA (padding, orientation, dedup weights), global-banded and local,
    with the unmodified abPOA's dumps."""
    return load_cases(REAL_PATH)


def load_real_meta(name):
    """Per block of a real case: (block id, padding, [(weight, strand flags, names)], final block graph text of the reference)."""
    z = np.load(REAL_PATH)
    out = []
    for line, gfa in zip(z[f"{name}/meta"], z[f"{name}/final_gfa"]):
        f = str(line).split("\t")
        seqs = []
        for rec in f[2:]:
            w, revs, names = rec.split(":", 2)
            seqs.append((int(w), [c == "1" for c in revs], names.split(",")))
        out.append((int(f[0]), int(f[1]), seqs, str(gfa)))
    return out


def load_cases(path=PATH):
    z = np.load(path)
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


# ---- full-size single blocks stored as per-section digests (tests/golden/make_deep_golden.py)
DEEP_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deep_golden.npz")
DEEP_CASES = {
    "deep_256x8kb": (dict(n_blocks=1, n_seqs=256, length=8000, divergence=0.02, seed=3001), dict()),
    "local_32x2kb": (dict(n_blocks=1, n_seqs=32, length=2000, divergence=0.02, seed=3002), dict(local=True)),
}
_SMALL = ("best_score", "n_cigar", "path_len", "cons_node")


def digest_dump(name, d):
    """{name/hdr, name/<small section>, name/sha/<section>} of a canonical dump (oracle.Dump)."""
    import hashlib
    from oracle.oracle import PD_FULL_LO
    out = {f"{name}/hdr": d.raw[:PD_FULL_LO].copy()}
    for k, v in d.sections().items():
        out[f"{name}/sha/{k}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(v, dtype=np.int32).tobytes()).digest(), dtype=np.uint8)
        if k in _SMALL:
            out[f"{name}/{k}"] = np.asarray(v, dtype=np.int32)
    return out


def deep_mismatches(name, d):
    """Sections of dump `d` whose digest differs from the unmodified abPOA's for DEEP_CASES[name] ([] = bit-exact)."""
    want = np.load(DEEP_PATH)
    got = digest_dump(name, d)
    return [k for k in got if k not in want.files or not np.array_equal(got[k], want[k])] + [k for k in want.files if k.startswith(name + "/") and k not in got]
