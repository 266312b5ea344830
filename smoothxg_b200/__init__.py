"""smoothxg_b200 -- B200-native partial-order-alignment engine behind smoothxg's per-block call site.

The product is smoothxg_b200/lib/libpoa_b200.so (C ABI in include/poa_b200.h, sm_100a CUDA in
smoothxg_b200/csrc).  This package holds the ctypes binding (engine.py), the seeded synthetic block
generator (synth.py) and the multi-GPU sharding helper (shard.py).  Nothing in this package loads the CPU checkers.
"""
__version__ = "0.1.0"
