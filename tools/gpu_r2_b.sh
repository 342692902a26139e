#!/bin/bash
# round-2 GPU call B (1 GPU): fixed tests, operand A/B, search experiments, ncu captures, the other workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dpr.py tests/test_gpu_search.py -m gpu -q --timeout=600 -p no:cacheprovider \
  -k "dpr_refresh or tier2 or non_finite or accumulation or caller_storage" > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/b_pytest.log
timeout 300 python tools/ab_operand.py 592x128 > gpurun_out/b_ab.log 2>&1; tail -1 gpurun_out/b_ab.log | cut -c1-900
timeout 900 python tools/exp_search_r2.py pace 8841823 18944,75776 > gpurun_out/b_pace.log 2>&1; tail -12 gpurun_out/b_pace.log | cut -c1-400
timeout 1500 python tools/exp_search_r2.py certify 8841823 18944 layernorm_clustered,iid,heavy_tail,near_duplicate,dpr > gpurun_out/b_certify.log 2>&1
tail -20 gpurun_out/b_certify.log | cut -c1-420
# ncu: coarse search in the tensor regime with the soft barrier (dram bytes), first two coarse launches after warm-up
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc05_gemm_kernel -s 1 -c 2 -o gpurun_out/prof_r2_search \
  python tools/exp_search_r2.py pace1 8841823 18944 > gpurun_out/b_ncu_search.log 2>&1
echo "ncu search rc=$?"
timeout 600 python bench.py --workload marco_doc_maxp --steps 3 --warmup 3 > gpurun_out/b_bench_maxp.json 2> gpurun_out/b_bench_maxp.err
echo "maxp rc=$?"; tail -c 600 gpurun_out/b_bench_maxp.json; tail -3 gpurun_out/b_bench_maxp.err
timeout 700 python bench.py --workload dpr --steps 3 --warmup 3 > gpurun_out/b_bench_dpr.json 2> gpurun_out/b_bench_dpr.err
echo "dpr rc=$?"; tail -c 600 gpurun_out/b_bench_dpr.json; tail -3 gpurun_out/b_bench_dpr.err
