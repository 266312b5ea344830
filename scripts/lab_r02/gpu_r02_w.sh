# round 2: ncu --set full of the leaner packed fill (one wave of 2368 blocks of configs[2] shape)
set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poa_b200_block -c 1 -o gpurun_out/r02w_full python bench.py --blocks 2368 --warps 1 --ctas-per-sm 16 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r02w_ncu_full.log 2>&1
tail -2 gpurun_out/r02w_ncu_full.log | cut -c1-200
