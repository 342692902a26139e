#!/bin/bash
# round-2 GPU call D (8 GPUs): ONE full 8,841,823-passage + 502,939-query refresh through the drop-in driver (north_star
# "Target"), in the roofline regime (every passage 128 real tokens) and with MS-MARCO-like lengths (variable-length tiles),
# then bench.py at N=8.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/d_gpus.txt; df -h /dev/shm /tmp >> gpurun_out/d_gpus.txt; nproc >> gpurun_out/d_gpus.txt
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > gpurun_out/d_clocks.csv &
SMI_PID=$!
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29551 tools/full_refresh.py --lengths full > gpurun_out/d_refresh_full.log 2>&1
echo "refresh full rc=$?"; tail -2 gpurun_out/d_refresh_full.log | cut -c1-2500
timeout 900 $RUN --master-port 29552 tools/full_refresh.py --lengths marco --tag marco > gpurun_out/d_refresh_marco.log 2>&1
echo "refresh marco rc=$?"; tail -2 gpurun_out/d_refresh_marco.log | cut -c1-2500
timeout 600 $RUN --master-port 29553 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/d_bench_n8.json 2> gpurun_out/d_bench_n8.err
echo "bench n8 rc=$?"; tail -c 800 gpurun_out/d_bench_n8.json; tail -3 gpurun_out/d_bench_n8.err
kill $SMI_PID 2>/dev/null
