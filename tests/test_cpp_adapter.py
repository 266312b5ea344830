"""The C++ host adapter (include/poa_b200_smooth.hpp: dedup -> encode -> GPU POA -> build_odgi-equivalent graph with
one path per name, reference src/smooth.cpp:133-627, :2442-2574).  CPU: the header compiles and links against the C ABI.
GPU (-m gpu): a block with duplicate and reverse-strand ranges gives the same graph as the ctypes path."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "smooth_adapter_test.cpp")
EXE = os.path.join(ROOT, "tests", "emu", "_build", "smooth_adapter_test")
LIBDIR = os.path.join(ROOT, "smoothxg_b200", "lib")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "include", "poa_b200_smooth.hpp"), os.path.join(ROOT, "include", "poa_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                               "-L", LIBDIR, "-lpoa_b200", f"-Wl,-rpath,{LIBDIR}"])
    return EXE


def test_adapter_header_compiles_and_links():
    from smoothxg_b200 import engine
    engine.load_library()  # the library must exist; no CPU fallback
    assert os.path.exists(build_exe())


def test_host_parameter_block_on_cpu():
    """build_params (smoothxg_b200/csrc/poa_host.hpp; reference src/smooth.cpp:256-297, abpoa_align.c:12-25,87-91): score matrix,
    packed constants of the 16-bit fill, and that the default-scoring instantiation of the fill is selected for smoothxg's
    defaults only (its literals are checked against the host-computed constants)."""
    exe = os.path.join(os.path.dirname(EXE), "params_test")
    src = os.path.join(ROOT, "tests", "cpp", "params_test.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "params ok", out.stdout + out.stderr


def test_dedup_and_encoding_on_cpu():
    """dedup_sequences = the XXH64 loop of src/smooth.cpp:217-241 (first-occurrence order, multiplicities, names and
    strands per group); encode_bases = ab_nt4_table."""
    exe = os.path.join(os.path.dirname(EXE), "dedup_test")
    src = os.path.join(ROOT, "tests", "cpp", "dedup_test.cpp")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", LIBDIR, "-lpoa_b200", f"-Wl,-rpath,{LIBDIR}"])
    inp = "a\t+\tACGT\nb\t-\tACGTN\nc\t+\tACGT\nd\t-\tACGT\ne\t+\tacgu\nf\t+\tACGTN\n"
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert out == ["3 a+ c+ d- 0123", "2 b- f+ 01234", "1 e+ 0123"]
    # adaptive_poa_preset: the reference compares the float estimate with double literals (0.95f < 0.95, 0.9f < 0.9)
    out = subprocess.run([exe, "0.995", "0.99", "0.98", "0.975", "0.95", "0.9500001", "0.9", "0.9000001", "0.7"], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert out == ["1 19 39 3 81 1", "1 19 39 3 81 1", "1 13 31 3 51 1", "1 9 16 2 41 1", "1 4 6 2 26 1", "1 7 11 2 33 1", "-", "1 4 6 2 26 1", "-"]


def test_extract_range_sequence_on_a_mock_graph():
    """poa_b200::extract_range_sequence == src/smooth.cpp:75-126,:177-214, quirks included (worked out by hand from the reference
    code; the patched smoothxg compares it with the reference's own strings on every real range, integration/)."""
    exe = os.path.join(os.path.dirname(EXE), "extract_test")
    src = os.path.join(ROOT, "tests", "cpp", "extract_test.cpp")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", LIBDIR, "-lpoa_b200", f"-Wl,-rpath,{LIBDIR}"])
    run = lambda *a: subprocess.run([exe, *map(str, a)], capture_output=True, text=True, check=True).stdout.strip()
    nodes = ["ACGT+", "GG+", "TTA+", "C+", "AAAA+"]
    # range = steps 2..3 (TTA C), padding 3: left flank starts AT step 2 (its last 3 bases: TTA), nothing more needed; right
    # flank from step 4: the LAST 3 bases of AAAA
    assert run(3, 2, 4, *nodes) == "TTATTACAAA 0"
    # padding 5: left takes TTA (3) then the last 2 of GG; step 0 is never visited; right runs off the path: AAAA + one N
    assert run(5, 2, 4, *nodes) == "TTAGGTTACAAAAN 0"
    # padding 9: left TTA, GG, then stops at the path's first step with 4 still missing -> NNNN in front
    assert run(9, 2, 4, *nodes) == "NNNNTTAGGTTACAAAANNNNN 0"
    # padding 0: the bare range
    assert run(0, 1, 3, *nodes) == "GGTTA 0"
    # mostly reverse-strand steps: the whole string is reverse-complemented
    assert run(0, 0, 3, "ACGT-", "GG-", "T+") == "AGGACGT 1"  # walk = ACGT (rc of the palindrome) CC T; 6 reverse bases > 1 forward


def _identity_exe():
    exe = os.path.join(os.path.dirname(EXE), "dedup_test")
    src = os.path.join(ROOT, "tests", "cpp", "dedup_test.cpp")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", LIBDIR, "-lpoa_b200", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def _identity_input():
    from tests.mash_cases import make_cases
    blocks = [seqs for _, k, seqs in make_cases(seed=31, n=10) if k == 17 and seqs]
    return blocks, "".join(f"{b} {s}\n" for b, seqs in enumerate(blocks) for s in seqs)


def test_identity_estimate_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _, inp = _identity_input()
    out = subprocess.run([_identity_exe(), "identity"], input=inp, capture_output=True, text=True)
    assert out.returncode == 3 and "identity estimate failed" in out.stdout


@pytest.mark.gpu
def test_identity_estimate_through_the_adapter():
    """poa_b200::estimate_block_identity (the batched src/smooth.cpp:1982-2023) == the oracle, bit for bit."""
    from oracle.mash import MashOracle
    blocks, inp = _identity_input()
    out = subprocess.run([_identity_exe(), "identity"], input=inp, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    got = [np.float32(x) for x in out.stdout.split()]
    ora = MashOracle()
    want = [ora.block(seqs, 17)[1] for seqs in blocks]
    assert got == [np.float32(-1.0) if w is None else w for w in want]


@pytest.mark.gpu
@pytest.mark.parametrize("padding,local,cons", [(0, 0, "Consensus_0"), (5, 0, "-"), (3, 1, "cons")])
def test_adapter_matches_ctypes_path(padding, local, cons):
    from smoothxg_b200 import engine as E
    from smoothxg_b200 import synth
    rng = np.random.default_rng(5)
    base = "".join("ACGT"[i] for i in rng.integers(0, 4, 160))
    def mut(s, k):
        s = list(s)
        for p in rng.integers(10, len(s) - 10, k):
            s[p] = "ACGT"[(("ACGT".index(s[p])) + 1) % 4]
        return "".join(s)
    a, b, c = mut(base, 3), mut(base, 5), base[:70] + "GGGTT" + base[70:]
    rows = [("p1", "+", base), ("p2", "-", a), ("p3", "+", base), ("p4", "+", b), ("p5", "-", a), ("p6", "+", c), ("p7", "+", "ACGTNNACGT" + base[:40])]
    inp = "".join(f"{n}\t{s}\t{q}\n" for n, s, q in rows)
    out = subprocess.run([build_exe(), str(padding), str(local), cons], input=inp, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    # the same block through the ctypes binding
    uniq, weights, names = [], [], []
    for n, s, q in rows:
        if q in uniq:
            k = uniq.index(q); weights[k] += 1; names[k].append((n, s))
        else:
            uniq.append(q); weights.append(1); names.append([(n, s)])
    batch = synth.PoaBatch.from_blocks([([E.encode_bases(q) for q in uniq], np.array(weights, dtype=np.int32))])
    eng = E.PoaEngine(device=0)
    res = eng.run_batch(batch, E.make_params(local=bool(local), out_msa=True, out_cons=cons != "-"))
    v = res.block(0)
    g = res.block_graph(0, padding, cons != "-")
    want = [f"dedup {len(uniq)} " + " ".join(str(w) for w in weights), f"nodes {len(g.node_id)}"]
    want += [f"S {i} {chr(b)}" for i, b in zip(g.node_id.tolist(), g.node_base)]
    want += [f"L {a_} {b_}" for a_, b_ in zip(g.edge_from.tolist(), g.edge_to.tolist())]
    for k, grp in enumerate(names):
        steps = g.path(k).tolist()
        for n, s in grp:
            st = [f"{x}-" for x in reversed(steps)] if s == "-" else [f"{x}+" for x in steps]
            want.append(" ".join([f"P {n}"] + st))
    if cons != "-":
        want.append(" ".join([f"P {cons}"] + [f"{x}+" for x in g.path(len(uniq)).tolist()]))
    want.append(f"msa {v.msa_rows} {v.msa_len}")
    lines = out.stdout.strip().split("\n")
    cut = next(i for i, ln in enumerate(lines) if ln.startswith("final "))
    assert lines[:cut] == want
    # final_graph_of (unchop + topological order), paths in the block's original order with duplicates and reverse strands
    res2 = eng.run_batch(batch, E.make_params(local=bool(local), out_msa=False, out_cons=cons != "-"))
    fg = res2.final_graph(0, padding, cons != "-")
    wantf = [f"final {len(fg.node_seq)}"] + [f"FS {k + 1} {sq}" for k, sq in enumerate(fg.node_seq)]
    wantf += [f"FL {a_} {b_}" for a_, b_ in zip(fg.edge_from.tolist(), fg.edge_to.tolist())]
    owner = {n: (k, s) for k, grp in enumerate(names) for n, s in grp}
    for n, _, _ in rows:
        k, s = owner[n]
        steps = fg.path(k).tolist()
        wantf.append(" ".join([f"FP {n}"] + ([f"{x}-" for x in reversed(steps)] if s == "-" else [f"{x}+" for x in steps])))
    if cons != "-":
        wantf.append(" ".join([f"FP {cons}"] + [f"{x}+" for x in fg.path(len(uniq)).tolist()]))
    assert lines[cut:] == wantf
    res.close(); res2.close(); eng.close()
