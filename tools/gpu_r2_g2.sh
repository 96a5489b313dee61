#!/bin/bash
# round 2, GPU call G2 (8 GPUs): MU bench at the driver's settings, in-stream profile, configs[3] column-sharded
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g2_bench_$N.json 2> gpurun_out/g2_bench_$N.err; echo "bench rc=$?" >> gpurun_out/g2_bench_$N.err
NMFGPU_PROFILE_ITERATION=1 timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/g2_profile_$N.json 2> gpurun_out/g2_profile_$N.err
timeout 900 $TR --master-port 29614 tools/run_config.py cfg4 --iters 20 > gpurun_out/g2_cfg4_$N.jsonl 2> gpurun_out/g2_cfg4_$N.err; echo "cfg4 rc=$?" >> gpurun_out/g2_cfg4_$N.err
grep "^{" gpurun_out/g2_bench_$N.json | cut -c1-330; tail -2 gpurun_out/g2_bench_$N.err; grep "iteration\]" gpurun_out/g2_profile_$N.err | head -9
cut -c1-250 gpurun_out/g2_cfg4_$N.jsonl; tail -2 gpurun_out/g2_cfg4_$N.err
