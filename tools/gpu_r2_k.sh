#!/bin/bash
# round-2 GPU call K (1 GPU): the new offline-evaluation GPU test; the full refresh on ONE GPU again with the final driver
# (whole-wave encoder passes, 4-wave query blocks, threaded post-processing), full-length and MS-MARCO-like lengths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -q --timeout=500 -p no:cacheprovider > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/k_pytest.log
timeout 900 python tools/full_refresh.py --lengths full > gpurun_out/k_refresh_full_n1.log 2>&1
echo "refresh full rc=$?"; tail -1 gpurun_out/k_refresh_full_n1.log | cut -c1-1500
timeout 900 python tools/full_refresh.py --lengths marco --tag marco > gpurun_out/k_refresh_marco_n1.log 2>&1
echo "refresh marco rc=$?"; tail -1 gpurun_out/k_refresh_marco_n1.log | cut -c1-1500
