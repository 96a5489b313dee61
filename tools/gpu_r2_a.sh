#!/bin/bash
# round 2, GPU call A: parity suite, bench at the driver's settings, in-stream iteration profile
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "not sharded" > gpurun_out/a_pytest_main.log 2>&1; echo "pytest main rc=$?" >> gpurun_out/a_pytest_main.log
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q --timeout 300 > gpurun_out/a_pytest_sharded.log 2>&1; echo "pytest sharded rc=$?" >> gpurun_out/a_pytest_sharded.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_20.json 2> gpurun_out/a_bench_20.err; echo "bench rc=$?" >> gpurun_out/a_bench_20.err
timeout 300 python tools/profile_iteration.py > gpurun_out/a_profile.txt 2>&1
tail -3 gpurun_out/a_pytest_main.log; tail -3 gpurun_out/a_pytest_sharded.log; cat gpurun_out/a_bench_20.json | cut -c1-1500; tail -5 gpurun_out/a_bench_20.err; cat gpurun_out/a_profile.txt
