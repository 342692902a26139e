#!/bin/bash
# round-2 GPU call C (2 GPUs): the 2-rank NCCL driver test, the variable-length encoder test, bench at N=2, a reduced-size
# full refresh through tools/full_refresh.py on 2 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_encoder.py -m gpu -q --timeout=800 -p no:cacheprovider -rs \
  -k "two_rank or varlen" > gpurun_out/c_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/c_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  tools/full_refresh.py --n_passages 600000 --n_queries 60000 --n_dev 2000 --lengths marco --tag small > gpurun_out/c_refresh_small.log 2>&1
echo "refresh small rc=$?"; tail -3 gpurun_out/c_refresh_small.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/c_bench_n2.json 2> gpurun_out/c_bench_n2.err
echo "bench n2 rc=$?"; tail -c 700 gpurun_out/c_bench_n2.json; tail -3 gpurun_out/c_bench_n2.err
