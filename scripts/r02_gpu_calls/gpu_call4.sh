#!/bin/bash
# r02 call 4 (2 GPUs): fused-exchange multi-GPU correctness (2-rank cases) + 2-GPU timings on plate(256)
mkdir -p gpurun_out
O=gpurun_out/r02_c4
( time timeout 1500 python -m pytest tests/test_multi_gpu.py -q -x ) > ${O}_pytest_multi.txt 2>&1
tail -25 ${O}_pytest_multi.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "timesteps_plate or operator_equals or dropin" > ${O}_pytest_single.txt 2>&1
tail -3 ${O}_pytest_single.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
EC3D_BENCH_GRID=256 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu > ${O}_b256_1.json 2> ${O}_b256_1.err
EC3D_BENCH_GRID=256 timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > ${O}_b256_2.json 2> ${O}_b256_2.err
EC3D_XFUSE=0 EC3D_BENCH_GRID=256 timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > ${O}_b256_2_unfused.json 2> ${O}_b256_2_unfused.err
EC3D_COMM=nccl EC3D_BENCH_GRID=256 timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > ${O}_b256_2_nccl.json 2> ${O}_b256_2_nccl.err
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu > ${O}_b512_2.json 2> ${O}_b512_2.err
for f in ${O}_b*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['config']['iters_per_step'], round(d['config']['ms_per_iteration'],4), {k:(round(v['ms'],4),round(v['frac'],3),v['slowest_rank']) for k,v in d['kernels'].items()})
except Exception as e: print('ERR',e)
PY
tail -3 ${f%.json}.err
done
