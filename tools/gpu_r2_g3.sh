#!/bin/bash
# round 2, GPU call G3 (N GPUs): sharded runs against the single-GPU run, MU bench at the driver's settings, in-stream profile
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tools/dist_check.py > gpurun_out/g3_dist_check_$N.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/g3_dist_check_$N.log
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g3_bench_$N.json 2> gpurun_out/g3_bench_$N.err; echo "bench rc=$?" >> gpurun_out/g3_bench_$N.err
NMFGPU_PROFILE_ITERATION=1 timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/g3_profile_$N.json 2> gpurun_out/g3_profile_$N.err
grep "rank 0" gpurun_out/g3_dist_check_$N.log | cut -c1-160; tail -1 gpurun_out/g3_dist_check_$N.log
grep "^{" gpurun_out/g3_bench_$N.json | cut -c1-330; tail -1 gpurun_out/g3_bench_$N.err; grep "iteration\]" gpurun_out/g3_profile_$N.err | head -8
