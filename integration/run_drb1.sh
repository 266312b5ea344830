#!/bin/bash
# On a GPU box: the reference's own ctest input (CMakeLists.txt:562-567) through the PATCHED smoothxg (integration/_build/smoothxg,
# built by integration/build_reference.sh poa_b200) with every abPOA block aligned by the B200 engine.
#   SMOOTHXG_POA_ENGINE=verify  each block is ALSO aligned by the unmodified abPOA inside the same process and the two graphs
#                               (node count, bases, out edges in final order with weights, consensus node ids) are compared;
#   the pipeline then continues with the GPU result, so smoothxg's own self-check (src/main.cpp:770-803: every path re-spelled
#   from the smoothed graph, exit(1) on any base difference) validates the laced graph built from GPU output.
# usage: bash integration/run_drb1.sh <out-dir> [threads]
set -x
OUT=${1:-gpurun_out/integration}; T=${2:-16}
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$OUT"; cd "$OUT"
cp "$HERE/_build/DRB1-3123.seqwish.gfa" drb1.gfa
for mode in "-A -Z" "-A"; do
  tag=$(echo $mode | tr -d ' -')
  SMOOTHXG_POA_ENGINE=verify "$HERE/_build/smoothxg" -t $T -g drb1.gfa -j 5k -e 5k -l 700,900,1100 -r 12 $mode -o smooth_$tag.gfa > run_$tag.log 2>&1
  echo "mode '$mode' verify rc=$?" | tee -a summary.txt
  grep "poa_b200" run_$tag.log | tee -a summary.txt
  grep -c "^S" smooth_$tag.gfa | sed "s/^/S-lines $tag: /" | tee -a summary.txt
done
# GPU only (no CPU abPOA at all), timing of the POA stage as smoothxg reports it
SMOOTHXG_POA_ENGINE=gpu "$HERE/_build/smoothxg" -t $T -g drb1.gfa -j 5k -e 5k -l 700,900,1100 -r 12 -A -Z -o smooth_gpu.gfa > run_gpu.log 2>&1
echo "gpu-only rc=$?" | tee -a summary.txt
grep "poa_b200" run_gpu.log | tee -a summary.txt
rm -f drb1.gfa.smooth.*.og
