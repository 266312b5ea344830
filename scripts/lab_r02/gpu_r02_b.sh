# round 2, call B: compact wire format parity on the GPU; ncu --set full of the single-chunk and two-chunk fills (one wave)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -5 gpurun_out/r02b_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --cpu-sample 256 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
cat gpurun_out/r02b_bench.json | cut -c1-3000
for v in ilp1_c16 ilp2_c12; do
  c=${v##*_c}
  POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:poa_b200_block -c 1 -o gpurun_out/r02b_full_$v python bench.py --blocks $((148*c)) --warps 1 --ctas-per-sm $c --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r02b_ncu_$v.log 2>&1
  ls -la gpurun_out/r02b_full_$v.ncu-rep
done
