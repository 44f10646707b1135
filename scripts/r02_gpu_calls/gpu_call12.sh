#!/bin/bash
# r02 call 12 (1 GPU): final state -- the full GPU test suite and a short bench line on plate(256)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c12_pytest.txt 2>&1
tail -6 gpurun_out/r02_c12_pytest.txt
EC3D_BENCH_GRID=256 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_c12_b256.json 2> gpurun_out/r02_c12_b256.err
tail -c 900 gpurun_out/r02_c12_b256.json
