#!/bin/bash
# round-2 GPU call M (2 GPUs): the final sharded_search (results merged in place) under NCCL: 2-rank driver test, bench N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=500 -p no:cacheprovider -rs > gpurun_out/m_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/m_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  tools/full_refresh.py --n_passages 400000 --n_queries 160001 --n_dev 2001 --lengths marco --tag small > gpurun_out/m_refresh_small.log 2>&1
echo "refresh small rc=$?"; tail -1 gpurun_out/m_refresh_small.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/m_bench_n2.json 2> gpurun_out/m_bench_n2.err
echo "bench n2 rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/m_bench_n2.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], sum(v for k,v in j['kernel_ms_per_step'].items() if k!='encoder_gemm'))"
