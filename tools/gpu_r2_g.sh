#!/bin/bash
# round-2 GPU call G (1 GPU): single-pass online softmax for L > 128 (parity + timing), encoder pass-size sweep, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_dpr.py tests/test_gpu_driver.py -m gpu -q --timeout=800 -p no:cacheprovider \
  > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/g_pytest.log
timeout 300 python tools/perf_encoder.py 148x512,296x256,592x128 > gpurun_out/g_perf.log 2>&1; cut -c1-420 gpurun_out/g_perf.log
for mt in 37888 18944 9472; do
  ANCE_B200_MAX_TOKENS=$mt timeout 300 python tools/perf_encoder.py 592x128 > gpurun_out/g_perf_mt$mt.log 2>&1
  echo "max_tokens $mt: $(cut -c1-330 gpurun_out/g_perf_mt$mt.log | tail -1)"
done
timeout 600 python bench.py --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/g_bench_n1.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['clocks'], {k:round(v,1) for k,v in j['kernel_ms_per_step'].items()})
PY
