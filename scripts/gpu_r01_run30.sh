# r30: compute-sanitizer memcheck over the v11 kernel (F recompute on the traceback path) and the identity-estimate kernels,
# then the final default bench line of the committed build.
set -x
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r30_memcheck_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/r30_memcheck_smoke.txt
tail -5 gpurun_out/r30_memcheck_smoke.txt
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_mash.py -m gpu -x -q -k "golden or fresh" > gpurun_out/r30_memcheck_mash.txt 2>&1; echo "rc=$?" >> gpurun_out/r30_memcheck_mash.txt
tail -5 gpurun_out/r30_memcheck_mash.txt
python bench.py > gpurun_out/r30_bench.json 2> gpurun_out/r30_bench.err
cat gpurun_out/r30_bench.json
