#!/bin/bash
# round 2, GPU call D: whole GPU suite after the pushN / sparsity / fp64-inverse changes, update-H panel widths, LS stability study
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/d_pytest.log; tail -25 gpurun_out/d_pytest.log
NMFGPU_UPDATE_H_COLS=16 python tools/profile_iteration.py 12544 10000 64 60 2>&1 | tail -9
NMFGPU_UPDATE_H_COLS=32 python tools/profile_iteration.py 12544 10000 64 60 2>&1 | grep "update H"
NMFGPU_UPDATE_H_COLS=64 python tools/profile_iteration.py 12544 10000 64 60 2>&1 | grep "update H"
python tools/ls_stability.py > gpurun_out/d_ls_inverse64.log 2>&1
NMFGPU_LS_SOLVE=inverse32 python tools/ls_stability.py > gpurun_out/d_ls_inverse32.log 2>&1
NMFGPU_LS_SOLVE=qr python tools/ls_stability.py > gpurun_out/d_ls_qr.log 2>&1
paste -d'\n' gpurun_out/d_ls_inverse64.log gpurun_out/d_ls_inverse32.log gpurun_out/d_ls_qr.log | cut -c1-125
