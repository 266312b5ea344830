"""Seeded block sets for the identity-estimate tests (shared by the golden generator, the CPU tests and the GPU tests)."""
import numpy as np


def _mutate(rng, s, d):
    out = []
    for c in s:
        x = rng.random()
        if x < d * 0.6:
            out.append("ACGT"[rng.integers(4)])
        elif x < d * 0.8:
            out.append(c); out.append("ACGT"[rng.integers(4)])
        elif x < d:
            pass
        else:
            out.append(c)
    return "".join(out)


def make_cases(seed=11, n=24):
    """-> list of (name, kmer, [strings]); covers every branch of src/smooth.cpp:1982-2023 and rkmh::compare."""
    rng = np.random.default_rng(seed)
    cases = []
    for t in range(n):
        L = int(rng.integers(150, 2600))
        S = int(rng.integers(2, 14))
        d = float(rng.choice([0.0, 0.001, 0.01, 0.02, 0.05, 0.12, 0.35]))
        k = int(rng.choice([17, 17, 17, 15, 11, 21, 32, 16]))
        base = "".join(rng.choice(list("ACGT"), L))
        if t % 4 == 1:
            base = base + base[: L // 3]                      # repeats: duplicate hashes inside a list
        seqs = [_mutate(rng, base, d) for _ in range(S)]
        if t % 5 == 0:
            p = len(seqs[0]) // 2
            seqs[0] = seqs[0][:p] + "NNNNN" + seqs[0][p + 5:]    # non-ACGT bytes: zero hashes
        if t % 6 == 2:
            seqs.append(seqs[-1])                                # identical strings: distance 0 (common == denom)
        if t % 6 == 3:
            seqs.append("".join(rng.choice(list("ACGT"), L)))    # unrelated string: common == 0
        if t % 7 == 4:
            seqs.insert(1, "ACGT" * 8)                           # shorter than 8*k: dropped
        if t % 7 == 5:
            seqs[1] = seqs[1][: max(8 * k, len(seqs[1]) // 3)]   # much shorter partner: min-size denominator
        cases.append((f"case{t}_L{L}_S{S}_d{d}_k{k}", k, seqs))
    cases.append(("one_long_one_short", 17, ["".join(rng.choice(list("ACGT"), 400)), "ACGTACGT"]))   # kept < 2
    cases.append(("empty_block", 17, []))
    cases.append(("single", 17, ["".join(rng.choice(list("ACGT"), 300))]))
    cases.append(("exactly_8k", 17, ["".join(rng.choice(list("ACGT"), 136)) for _ in range(3)]))
    cases.append(("all_N", 17, ["N" * 200, "N" * 180, "".join(rng.choice(list("ACGT"), 300))]))
    poly = "A" * 500
    cases.append(("homopolymer", 17, [poly, poly[:400], poly[:300] + "C" + poly[:100]]))               # one value, huge multiplicity
    return cases
