#!/bin/bash
# round 2: ncu evidence of the final build -- launch list of bench.py, full capture of the two tensor-core products and of the update kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_stream_gemm|update_._fused|reduce_push" -c 12 -o gpurun_out/r02_iteration_kernels python tools/run_products.py --iters 4 --reps 1 > gpurun_out/r02_iteration_ncu.log 2>&1
tail -2 gpurun_out/r02_iteration_ncu.log; wc -l gpurun_out/r02_bench_launches.csv
