#!/bin/bash
# round-2 GPU call F (2 GPUs): 2-rank NCCL driver test (byte-identical to 1 rank), reduced-size full refresh on 2 ranks with
# a ragged last query block, bench at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_search.py -m gpu -q --timeout=800 -p no:cacheprovider -rs \
  -k "two_rank or uncertified or centering" > gpurun_out/f_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/f_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  tools/full_refresh.py --n_passages 600000 --n_queries 60001 --n_dev 2001 --lengths marco --tag small > gpurun_out/f_refresh_small.log 2>&1
echo "refresh small rc=$?"; tail -1 gpurun_out/f_refresh_small.log | cut -c1-1700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/f_bench_n2.json 2> gpurun_out/f_bench_n2.err
echo "bench n2 rc=$?"; tail -c 500 gpurun_out/f_bench_n2.json; tail -2 gpurun_out/f_bench_n2.err
