#!/bin/bash
# round 2, GPU call C: the two products on the shape of an 8-GPU row block (12544 x 10000), single GPU
mkdir -p gpurun_out
python tools/run_products.py --m 12544 --n 10000 --reps 6 > gpurun_out/c_products_shard.log 2>&1
python tools/profile_iteration.py 12544 10000 64 60 > gpurun_out/c_profile_shard.log 2>&1
python tools/run_products.py --m 50176 --n 10000 --reps 4 > gpurun_out/c_products_shard2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_stream_gemm -c 4 -o gpurun_out/c_tc_shard python tools/run_products.py --m 12544 --n 10000 --reps 2 > gpurun_out/c_ncu.log 2>&1
cat gpurun_out/c_products_shard.log gpurun_out/c_profile_shard.log gpurun_out/c_products_shard2.log; tail -3 gpurun_out/c_ncu.log
