#!/bin/bash
# Build the reference smoothxg from a WRITABLE COPY (the reference's CMake writes into its source tree), with the edits this
# container needs (no libzstd-dev / libjemalloc-dev: SURVEY.md 8c) and, optionally, one of the patches in this directory:
#   integration/build_reference.sh harvest    -> -DPOA_DEBUG=ON + harvest.patch   (fixture harvest: tests/golden/make_real_golden.py)
#   integration/build_reference.sh poa_b200   -> smoothxg_poa_b200.patch, linked against smoothxg_b200/lib/libpoa_b200.so
# Result: /tmp/sxg/bin/smoothxg (about 15 minutes the first time: odgi dominates; seconds afterwards).
# Never run smoothxg on a GFA that lies under /root/reference: it writes <input>.smooth.<i>.og next to its input.
set -e
MODE=${1:-poa_b200}
REPO=$(cd "$(dirname "$0")/.." && pwd)
if [ ! -d /tmp/sxg ]; then
  cp -r /root/reference /tmp/sxg
  (cd /tmp/sxg && patch -p0 CMakeLists.txt < "$REPO/integration/CMakeLists.patch")
  cp /tmp/sxg/src/smooth.cpp /tmp/sxg/src/smooth.cpp.orig
  mkdir -p /tmp/sxg_stub && echo 'void __jemalloc_stub(void){}' > /tmp/sxg_stub/s.c && /usr/bin/gcc -shared -fPIC -o /tmp/sxg_stub/libjemalloc.so /tmp/sxg_stub/s.c
fi
cd /tmp/sxg
cp src/smooth.cpp.orig src/smooth.cpp
if [ "$MODE" = harvest ]; then patch src/smooth.cpp < "$REPO/integration/harvest.patch"; DBG=ON; else patch src/smooth.cpp < "$REPO/integration/smoothxg_poa_b200.patch"; DBG=OFF; fi
export CC=/usr/bin/gcc CXX=/usr/bin/g++ LIBRARY_PATH=/tmp/sxg_stub LD_LIBRARY_PATH=/tmp/sxg_stub
cmake -S . -B build -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ -DCMAKE_BUILD_TYPE=Release -DEXTRA_FLAGS="-O3 -march=native" -DPOA_DEBUG=$DBG
cmake --build build -j 6
if [ "$MODE" = poa_b200 ]; then
  mkdir -p "$REPO/integration/_build"
  cp bin/smoothxg "$REPO/integration/_build/smoothxg"   # git-ignored; travels to the GPU box (RUNPATH $ORIGIN/../../smoothxg_b200/lib)
  cp test/data/DRB1-3123.fa.gz.pggb-s3000-p70-n10-a70-K16-k8-w10000-j5000-e5000.seqwish.gfa "$REPO/integration/_build/DRB1-3123.seqwish.gfa"
fi
