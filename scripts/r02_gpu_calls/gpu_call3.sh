#!/bin/bash
# r02 call 3 (1 GPU): full GPU test suite incl. the parity-at-scale tests, fork/join SpMV timings, decks
mkdir -p gpurun_out
O=gpurun_out/r02_c3
( time timeout 2400 python -m pytest tests -m gpu -q --durations=15 ) > ${O}_pytest.txt 2>&1
tail -40 ${O}_pytest.txt
for N in 512 256; do
  python scripts/spmv_bench.py $N 0,1,5                    >> ${O}_kern.jsonl 2>&1
  EC3D_FORK=0 python scripts/spmv_bench.py $N 0,1,5        >> ${O}_kern.jsonl 2>&1
  EC3D_CCPS=1 python scripts/spmv_bench.py $N 0,1,5        >> ${O}_kern.jsonl 2>&1
done
EC3D_FUSE_S=0 python scripts/spmv_bench.py 256 0,1,2,5   >> ${O}_kern.jsonl 2>&1
cat ${O}_kern.jsonl
python scripts/deck_bench.py 20 > ${O}_decks.jsonl 2>&1
cat ${O}_decks.jsonl
python scripts/dropin_bench.py 128 > ${O}_dropin.jsonl 2>&1
cat ${O}_dropin.jsonl
