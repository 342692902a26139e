#!/bin/bash
# round-2 GPU call I (1 GPU): what the driver runs at round end — whole GPU suite, smoke(), default bench (both arms)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -rs > gpurun_out/i_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/i_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/i_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/i_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/i_bench_reference.json 2> gpurun_out/i_bench_reference.err
echo "reference arm rc=$?"; cut -c1-400 gpurun_out/i_bench_reference.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/i_bench_n1.json 2> gpurun_out/i_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/i_bench_n1.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], j['clocks'])
print({k:round(v,1) for k,v in j['kernel_ms_per_step'].items()}); print(j['stages']['passages_per_s_marco_like_lengths'], j['stages']['queries_topk_per_s'], j['cpu_baseline']['value'])
PY
