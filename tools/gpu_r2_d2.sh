#!/bin/bash
# round-2 GPU call D2 (8 GPUs): the full 8,841,823 + 502,939 refresh again after the fixes found by the first run
# (results padded instead of queries, NCCL warmed at start-up, whole-wave encoder passes, threaded post-processing)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > gpurun_out/d2_clocks.csv &
SMI_PID=$!
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29551 tools/full_refresh.py --lengths full > gpurun_out/d2_refresh_full.log 2>&1
echo "refresh full rc=$?"; tail -1 gpurun_out/d2_refresh_full.log | cut -c1-2300
timeout 600 $RUN --master-port 29552 tools/full_refresh.py --lengths marco --tag marco > gpurun_out/d2_refresh_marco.log 2>&1
echo "refresh marco rc=$?"; tail -1 gpurun_out/d2_refresh_marco.log | cut -c1-2300
kill $SMI_PID 2>/dev/null
