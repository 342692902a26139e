#!/bin/bash
# round-2 GPU call L (1 GPU): packed-register LayerNorm variant (bit-identity test + alternating A/B), data.py rewrite under the GPU driver tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_driver.py tests/test_trainer_side.py -m gpu -q --timeout=500 -p no:cacheprovider \
  -k "rows_per_warp or refresh_end_to_end or inference or forward_loss or golden" > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
for rep in 1 2 3; do for v in 2 3; do
  ANCE_LN_ROWS=$v timeout 200 python tools/perf_encoder.py 592x128 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ln_rows=$v', 'ms', round(j['ms'],3), 'norm_ms', round(j['norm_ms'],4), 'gemm', round(j['gemm_ms'],3), 'attn', round(j['attn_ms'],3))"
done; done
