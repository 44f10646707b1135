#!/bin/bash
# r02 call 6 (8 GPUs): 4- and 8-rank parity, strong scaling plate(512) at 8 / 4 ranks, plate(768) on 8 ranks
mkdir -p gpurun_out
O=gpurun_out/r02_c6
( time EC3D_TEST_MIN_RANKS=4 timeout 1200 python -m pytest tests/test_multi_gpu.py -q -x ) > ${O}_pytest_multi.txt 2>&1
tail -8 ${O}_pytest_multi.txt
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552"
timeout 900 $TR8 bench.py --gpus 8 --steps 4 --warmup 2 --no-cpu > ${O}_b512_8.json 2> ${O}_b512_8.err
EC3D_XFUSE=0 timeout 900 $TR8 bench.py --gpus 8 --steps 4 --warmup 2 --no-cpu > ${O}_b512_8_unfused.json 2> ${O}_b512_8_unfused.err
timeout 900 $TR4 bench.py --gpus 4 --steps 4 --warmup 2 --no-cpu > ${O}_b512_4.json 2> ${O}_b512_4.err
EC3D_BENCH_GRID=768 timeout 1500 $TR8 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu > ${O}_b768_8.json 2> ${O}_b768_8.err
for f in ${O}_b*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['e2e']['value'], d['config']['iters_per_step'], round(d['config']['ms_per_iteration'],4), {k:(round(v['ms'],4),round(v['frac'],3),v['slowest_rank']) for k,v in d['kernels'].items()}, d['clocks'])
except Exception as e: print('ERR',e)
PY
tail -2 ${f%.json}.err | cut -c1-300
done
