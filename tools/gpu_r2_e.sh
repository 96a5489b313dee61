#!/bin/bash
# round 2, GPU call E: whole GPU suite after folding the signals into the neighbouring kernels, iteration profile, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/e_pytest.log; tail -25 gpurun_out/e_pytest.log
python tools/profile_iteration.py 2>&1 | tail -9
python tools/profile_iteration.py 12544 10000 64 60 2>&1 | tail -9
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_20.json 2> gpurun_out/e_bench_20.err; cut -c1-400 gpurun_out/e_bench_20.json; tail -3 gpurun_out/e_bench_20.err
