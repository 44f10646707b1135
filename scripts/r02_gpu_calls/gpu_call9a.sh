#!/bin/bash
# r02 call 9a (1 GPU): the full GPU test suite (incl. plate(128) / plate(256) bridge hashes) and plate(768) on one GPU
mkdir -p gpurun_out
O=gpurun_out/r02_c9a
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > ${O}_pytest.txt 2>&1
tail -16 ${O}_pytest.txt
EC3D_BENCH_GRID=768 timeout 1500 python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu > ${O}_b768_1.json 2> ${O}_b768_1.err
tail -c 1800 ${O}_b768_1.json; tail -n 3 ${O}_b768_1.err
