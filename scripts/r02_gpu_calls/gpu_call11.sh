#!/bin/bash
# r02 call 11 (2 GPUs): plate(512) on 2 ranks with the final build (the 2-GPU line of the scaling table)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 900 $TR bench.py --gpus 2 --steps 4 --warmup 2 --no-cpu > gpurun_out/r02_c11_b512_2.json 2> gpurun_out/r02_c11_b512_2.err
tail -c 600 gpurun_out/r02_c11_b512_2.json
