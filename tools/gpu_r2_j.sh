#!/bin/bash
# round-2 GPU call J (8 GPUs): the MaxP (configs[3]) and DPR (configs[4]) workloads at N=8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $RUN --master-port 29561 bench.py --gpus 8 --workload marco_doc_maxp --steps 3 --warmup 3 > gpurun_out/j_bench_maxp_n8.json 2> gpurun_out/j_bench_maxp_n8.err
echo "maxp n8 rc=$?"; tail -c 400 gpurun_out/j_bench_maxp_n8.json; tail -2 gpurun_out/j_bench_maxp_n8.err | cut -c1-300
timeout 500 $RUN --master-port 29562 bench.py --gpus 8 --workload dpr --steps 3 --warmup 3 > gpurun_out/j_bench_dpr_n8.json 2> gpurun_out/j_bench_dpr_n8.err
echo "dpr n8 rc=$?"; tail -c 400 gpurun_out/j_bench_dpr_n8.json; tail -2 gpurun_out/j_bench_dpr_n8.err | cut -c1-300
