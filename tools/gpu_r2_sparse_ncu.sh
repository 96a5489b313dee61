#!/bin/bash
# round 2: ncu capture of the two SpMM gathers of the sparse execution (a quarter of configs[4]: 250 000 x 100 000, nnz 2.5e7, k = 100)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none -k regex:spmm_gather -s 4 -c 2 -o gpurun_out/r02_spmm python tools/run_config.py cfg5 --m 250000 --iters 3 > gpurun_out/r02_spmm_ncu.log 2>&1
tail -3 gpurun_out/r02_spmm_ncu.log
