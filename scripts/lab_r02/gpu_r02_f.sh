# round 2, call F (one GPU): the patched smoothxg on the reference's own test input, every abPOA block on the GPU, verify mode
set -x
mkdir -p gpurun_out
bash integration/run_drb1.sh gpurun_out/integration 16 > gpurun_out/r02f_integration.log 2>&1
cat gpurun_out/integration/summary.txt
tail -3 gpurun_out/integration/run_AZ.log | cut -c1-300
