#!/bin/bash
# round 2, final 1-GPU validation: whole GPU suite, smoke, bench at the driver's settings (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | grep -v "^$" | tail -5 > gpurun_out/z_pytest.log; cat gpurun_out/z_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench_20.json 2> gpurun_out/z_bench_20.err; cut -c1-260 gpurun_out/z_bench_20.json; tail -2 gpurun_out/z_bench_20.err
