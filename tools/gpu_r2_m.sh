#!/bin/bash
# N-GPU quick check: sharded correctness (dist_check), bench at the driver's settings, in-stream profile (40 iterations)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/dist_check.py > gpurun_out/m_dist_check_$N.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/m_dist_check_$N.log
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m_bench_$N.json 2> gpurun_out/m_bench_$N.err; echo "bench rc=$?" >> gpurun_out/m_bench_$N.err
NMFGPU_PROFILE_ITERATION=1 timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/m_profile_$N.json 2> gpurun_out/m_profile_$N.err
grep -c " ok" gpurun_out/m_dist_check_$N.log; grep -c MISMATCH gpurun_out/m_dist_check_$N.log; tail -1 gpurun_out/m_dist_check_$N.log
grep "^{" gpurun_out/m_bench_$N.json | cut -c1-330; tail -1 gpurun_out/m_bench_$N.err; grep "iteration\]" gpurun_out/m_profile_$N.err | head -8
