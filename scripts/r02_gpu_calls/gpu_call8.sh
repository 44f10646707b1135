#!/bin/bash
# r02 call 8 (8 GPUs): fused exchange after the one-fence-per-block fix vs stand-alone exchange kernels
mkdir -p gpurun_out
O=gpurun_out/r02_c8
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562"
timeout 900 $TR8 bench.py --gpus 8 --steps 4 --warmup 2 --no-cpu > ${O}_b512_8.json 2> ${O}_b512_8.err
EC3D_XFUSE=0 timeout 900 $TR8 bench.py --gpus 8 --steps 4 --warmup 2 --no-cpu > ${O}_b512_8_unfused.json 2> ${O}_b512_8_unfused.err
timeout 900 $TR4 bench.py --gpus 4 --steps 4 --warmup 2 --no-cpu > ${O}_b512_4.json 2> ${O}_b512_4.err
( EC3D_TEST_MIN_RANKS=8 timeout 600 python -m pytest tests/test_multi_gpu.py -q -x ) > ${O}_pytest_fused.txt 2>&1
( EC3D_XFUSE=0 EC3D_TEST_MIN_RANKS=8 timeout 600 python -m pytest tests/test_multi_gpu.py -q -x ) > ${O}_pytest_unfused.txt 2>&1
tail -2 ${O}_pytest_fused.txt ${O}_pytest_unfused.txt
for f in ${O}_b*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['e2e']['value'], d['config']['iters_per_step'], round(d['config']['ms_per_iteration'],4))
    for k,v in d['kernels'].items(): print('   ',k, round(v['ms'],4),round(v['frac'],3),v['slowest_rank'], v.get('ms_by_rank'))
    print('   ', d.get('owned_unknowns_by_rank'))
except Exception as e: print('ERR',e)
PY
tail -2 ${f%.json}.err | cut -c1-300
done
