#!/bin/bash
# r02 call 7 (1 GPU): tail-split item plan on plate(256) / plate(512), decks with the short-item default
mkdir -p gpurun_out
O=gpurun_out/r02_c7
for N in 256 512; do
  python scripts/spmv_bench.py $N 0,1,5 >> ${O}_kern.jsonl 2>&1
  EC3D_TAIL=1 python scripts/spmv_bench.py $N 0,1,5 >> ${O}_kern.jsonl 2>&1
done
EC3D_TAIL=1 EC3D_ZC=24 python scripts/spmv_bench.py 256 0,1,5 >> ${O}_kern.jsonl 2>&1
EC3D_TAIL=1 EC3D_ZC=16 python scripts/spmv_bench.py 256 0,1,5 >> ${O}_kern.jsonl 2>&1
cat ${O}_kern.jsonl
python scripts/deck_bench.py 20 > ${O}_decks.jsonl 2>&1
cat ${O}_decks.jsonl
EC3D_TAIL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "operator or timesteps_plate" > ${O}_pytest.txt 2>&1
tail -3 ${O}_pytest.txt
