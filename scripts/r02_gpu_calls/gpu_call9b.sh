#!/bin/bash
# r02 call 9b (1 GPU): the driver's own command line -- python bench.py --gpus 1 --steps 20 --warmup 5
mkdir -p gpurun_out
O=gpurun_out/r02_c9b
( time timeout 1700 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench.json 2> ${O}_bench.err
tail -c 3000 ${O}_bench.json; tail -n 5 ${O}_bench.err
