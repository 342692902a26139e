#!/bin/bash
# round-2 GPU call N (2 GPUs): bench at N=2 with the search of slice i finished after slice i+1 is enqueued
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/n_bench_n2.json 2> gpurun_out/n_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/n_bench_n2.err | cut -c1-300; python -c "
import json; j=json.loads(open('gpurun_out/n_bench_n2.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], sum(v for k,v in j['kernel_ms_per_step'].items() if k!='encoder_gemm'), j['stages']['search_stats'])"
timeout 300 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/n_bench_n1.json 2> gpurun_out/n_bench_n1.err
echo "bench n1 rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/n_bench_n1.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], sum(v for k,v in j['kernel_ms_per_step'].items() if k!='encoder_gemm'))"
