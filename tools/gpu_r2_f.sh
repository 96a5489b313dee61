#!/bin/bash
# round 2, GPU call F (1 GPU): configs[3] and configs[4] on record, bench with the clock window
mkdir -p gpurun_out
timeout 900 python tools/run_config.py cfg4 --iters 20 > gpurun_out/f_cfg4_1gpu.jsonl 2> gpurun_out/f_cfg4_1gpu.err; cat gpurun_out/f_cfg4_1gpu.jsonl; tail -3 gpurun_out/f_cfg4_1gpu.err
timeout 900 python tools/run_config.py cfg5 --iters 20 > gpurun_out/f_cfg5_1gpu.jsonl 2> gpurun_out/f_cfg5_1gpu.err; cat gpurun_out/f_cfg5_1gpu.jsonl; tail -3 gpurun_out/f_cfg5_1gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_20.json 2> gpurun_out/f_bench_20.err; cat gpurun_out/f_bench_20.json; tail -3 gpurun_out/f_bench_20.err
