#!/bin/bash
# r02 call 5 (1 GPU): tests, ZC sweep, ncu evidence (launch list + full captures), deck kernel timings, Jacobi
mkdir -p gpurun_out
O=gpurun_out/r02_c5
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
python scripts/spmv_bench.py 512 > ${O}_kern.jsonl 2>&1
python scripts/spmv_bench.py 256 >> ${O}_kern.jsonl 2>&1
for zc in 16 24 32 48 64; do EC3D_ZC=$zc python scripts/spmv_bench.py 256 0,1,5 >> ${O}_kern.jsonl 2>&1; done
for zc in 24 32 64; do EC3D_ZC=$zc python scripts/spmv_bench.py 512 0,1,5 >> ${O}_kern.jsonl 2>&1; done
cat ${O}_kern.jsonl
python scripts/deck_bench.py 10 > ${O}_decks.jsonl 2>&1
EC3D_FORK=0 python scripts/deck_bench.py 10 >> ${O}_decks.jsonl 2>&1
cat ${O}_decks.jsonl
python scripts/jacobi_compare.py 256 6 > ${O}_jacobi.json 2>&1
cat ${O}_jacobi.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_tma|k_xr_update_tma|k_p_update_tma" -c 6 \
   -o gpurun_out/r02_kernels_512 -f python scripts/spmv_bench.py 512 0,1,3,4 0 1 > ${O}_ncu_full.log 2>&1
tail -3 ${O}_ncu_full.log
EC3D_BENCH_GRID=256 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_plate256.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu > ${O}_ncu_launches.log 2>&1
tail -2 ${O}_ncu_launches.log | cut -c1-300
