#!/bin/bash
# round-2 GPU call A (1 GPU): parity tests, bench, encoder per-class timings for both storage formats
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
python -c "import faiss; print('faiss', faiss.__version__)" > gpurun_out/a_faiss_probe.txt 2>&1
nproc >> gpurun_out/a_faiss_probe.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest_gpu.log
tail -5 gpurun_out/a_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/a_bench_n1.json; tail -5 gpurun_out/a_bench_n1.err
timeout 300 python tools/perf_encoder.py 592x128,148x512,296x256,1184x64 > gpurun_out/a_perf_fp16.log 2>&1
ANCE_B200_ENCODER_OPERAND=bf16 timeout 300 python tools/perf_encoder.py 592x128 > gpurun_out/a_perf_bf16.log 2>&1
tail -2 gpurun_out/a_perf_fp16.log gpurun_out/a_perf_bf16.log
