#!/bin/bash
# r02 call 10 (1 GPU): ncu --set full of the final kernels (default item length 32) on plate(512)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_tma|k_xr_update_tma|k_p_update_tma" -c 6 \
   -o gpurun_out/r02_kernels_512_final -f python scripts/spmv_bench.py 512 0,1,3,4 0 1 > gpurun_out/r02_c10_ncu.log 2>&1
tail -3 gpurun_out/r02_c10_ncu.log
