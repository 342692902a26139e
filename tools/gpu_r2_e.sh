#!/bin/bash
# round-2 GPU call E (1 GPU): whole GPU suite, small refresh (certification with a random-weight encoder), ncu captures
# (coarse search with the soft barrier; encoder GEMMs: DRAM traffic per launch; launch list of a bench slice), bench N=1,
# full 8.84M + 503k refresh on ONE GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -rs > gpurun_out/e_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/e_pytest_gpu.log
timeout 600 python tools/full_refresh.py --n_passages 600000 --n_queries 60000 --n_dev 2000 --lengths marco --tag small > gpurun_out/e_refresh_small.log 2>&1
echo "refresh small rc=$?"; tail -1 gpurun_out/e_refresh_small.log | cut -c1-1800
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc05_gemm_kernel -s 1 -c 1 -o gpurun_out/prof_r2_search \
  python tools/exp_search_r2.py pace1 8841823 18944 > gpurun_out/e_ncu_search.log 2>&1
echo "ncu search rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc05_gemm_kernel -s 100 -c 4 -o gpurun_out/prof_r2_gemm \
  python tools/perf_encoder.py 592x128 > gpurun_out/e_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 3200 --csv --log-file gpurun_out/e_launches.csv \
  python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/e_bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 700 python bench.py --steps 6 --warmup 3 > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
echo "bench rc=$?"; tail -c 900 gpurun_out/e_bench_n1.json; tail -3 gpurun_out/e_bench_n1.err
timeout 1200 python tools/full_refresh.py --lengths full > gpurun_out/e_refresh_full_n1.log 2>&1
echo "refresh n1 rc=$?"; tail -1 gpurun_out/e_refresh_full_n1.log | cut -c1-2200
