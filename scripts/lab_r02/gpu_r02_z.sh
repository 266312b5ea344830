# round 2: register-allocation hint of the single-warp kernel after the row-loop diet (launch bounds 13 / 14 / 15 resident blocks; all 128 registers)
set -x
bash scripts/gpu_variants.sh r02z
