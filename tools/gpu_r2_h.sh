#!/bin/bash
# round 2, GPU call H (1 GPU): launch lists of the LS family at configs[3], ncu of the update kernels, compute-sanitizer
mkdir -p gpurun_out
for algo in gdcls ahcls; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/h_cfg4_${algo}_launches.csv python tools/run_config.py cfg4 --algo $algo --iters 2 --init random > gpurun_out/h_cfg4_${algo}_ncu.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"update_._fused|reduce_push" -c 6 -o gpurun_out/h_update_kernels python tools/run_products.py --iters 4 --reps 1 > gpurun_out/h_update_ncu.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "test_v_sized_products and (1000-500-10 or 777-333-64 or 260-130-7)" > gpurun_out/h_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "test_v_sized_products and (1000-500-10 or 260-130-7) and auto" > gpurun_out/h_sanitizer_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_sharded_gpu.py -m gpu -q -x -k "test_mu_shards_match_single_gpu and rowblocks and 2500" > gpurun_out/h_sanitizer_memcheck_sharded.log 2>&1
tail -4 gpurun_out/h_sanitizer_memcheck.log gpurun_out/h_sanitizer_racecheck.log gpurun_out/h_sanitizer_memcheck_sharded.log; tail -2 gpurun_out/h_update_ncu.log
