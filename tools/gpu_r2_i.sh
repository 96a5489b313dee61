#!/bin/bash
# round 2, GPU call I (1 GPU): suite after the multi-panel W update / per-block fence / fast k x k inverse; profiles; cfg4; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/i_pytest.log; tail -12 gpurun_out/i_pytest.log
python tools/profile_iteration.py 2>&1 | tail -8
NMFGPU_UPDATE_H_COLS=32 python tools/profile_iteration.py 2>&1 | grep "update H"
python tools/profile_iteration.py 12544 10000 64 60 2>&1 | tail -8
timeout 900 python tools/run_config.py cfg4 --iters 20 > gpurun_out/i_cfg4_1gpu.jsonl 2> gpurun_out/i_cfg4_1gpu.err; cut -c1-260 gpurun_out/i_cfg4_1gpu.jsonl; tail -3 gpurun_out/i_cfg4_1gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/i_bench_20.json 2> gpurun_out/i_bench_20.err; cut -c1-300 gpurun_out/i_bench_20.json; tail -3 gpurun_out/i_bench_20.err
timeout 600 python bench.py > gpurun_out/i_bench_200.json 2> gpurun_out/i_bench_200.err; cut -c1-300 gpurun_out/i_bench_200.json; tail -3 gpurun_out/i_bench_200.err
