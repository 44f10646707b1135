#!/bin/bash
# r02 call 2: correctness of the fused s-update SpMV / TMA-ring BLAS-1 kernels, variant timings
mkdir -p gpurun_out
O=gpurun_out/r02_c2
timeout 600 python scripts/debug_step.py > ${O}_debug.txt 2>&1
tail -12 ${O}_debug.txt
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/debug_step.py > ${O}_memcheck.txt 2>&1
tail -5 ${O}_memcheck.txt
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.txt 2>&1
tail -15 ${O}_pytest.txt
for N in 512 256; do
  python scripts/spmv_bench.py $N                         >> ${O}_kern.jsonl 2>&1
  EC3D_CCPS=1 python scripts/spmv_bench.py $N             >> ${O}_kern.jsonl 2>&1
  EC3D_RING=0 python scripts/spmv_bench.py $N 3,4         >> ${O}_kern.jsonl 2>&1
  EC3D_RING=0 EC3D_VPAD=0 python scripts/spmv_bench.py $N 3,4  >> ${O}_kern.jsonl 2>&1
  EC3D_FUSE_S=0 python scripts/spmv_bench.py $N 0,1,2     >> ${O}_kern.jsonl 2>&1
  EC3D_SEGPAD=2064 python scripts/spmv_bench.py $N 0,1    >> ${O}_kern.jsonl 2>&1
  EC3D_SEGPAD=131584 python scripts/spmv_bench.py $N 0,1  >> ${O}_kern.jsonl 2>&1
  EC3D_SEGPAD=528 EC3D_CCPS=1 python scripts/spmv_bench.py $N 0,1  >> ${O}_kern.jsonl 2>&1
done
cat ${O}_kern.jsonl
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > ${O}_bench.json 2> ${O}_bench.err
tail -c 1500 ${O}_bench.json
