#!/bin/bash
# r02 call 1: streaming probe, ncu of the BLAS-1 kernels at 512^3, single-pass bench check, GPU tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw --format=csv > gpurun_out/r02_c1_smi.txt
./scripts/stream_probe > gpurun_out/r02_stream_probe.txt 2>&1
python scripts/spmv_bench.py 512 > gpurun_out/r02_c1_kern512.json 2>&1
python scripts/spmv_bench.py 256 > gpurun_out/r02_c1_kern256.json 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_s_update|k_xr_update|k_p_update" -c 3 \
   -o gpurun_out/r02_blas1_512 -f python scripts/spmv_bench.py 512 2,3,4 0 1 > gpurun_out/r02_c1_ncu.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c1_pytest.txt 2>&1
tail -5 gpurun_out/r02_c1_pytest.txt
cat gpurun_out/r02_stream_probe.txt
