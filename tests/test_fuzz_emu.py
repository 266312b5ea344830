"""Bounded run of scripts/fuzz_emu.py in the CPU suite: random scoring parameters (convex / affine / linear), modes and
small random blocks through the emulated device code (1 lane and 32 lock-step lanes) against the oracle, and the oracle
against the unmodified abPOA where oracle/_ref exists.  `python scripts/fuzz_emu.py 1000 <seed>` is the long form."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_emulated_device_logic(tmp_path):
    # the emulation libraries are built by tests/test_emu_parity.py's fixtures; build them here too if this file runs alone
    from tests import test_emu_parity as T
    for out, flags in ((T.OUT, []), (T.OUT32, ["-DPOA_EMU_LANES=32"])):
        deps = [T.SRC] + [os.path.join(ROOT, "smoothxg_b200", "csrc", f) for f in ("poa_core.cuh", "poa_fill16.cuh", "poa_host.hpp", "poa_wire.hpp")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.check_call(["/usr/bin/g++", "-O1", "-fPIC", "-shared", "-std=c++17", *flags, "-I/usr/local/cuda/include", "-o", out, T.SRC])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_emu.py"), "16", "20261017"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "fuzz ok: 16 cases" in r.stdout
