#!/bin/bash
# round-2 GPU call H (1 GPU): ncu --set full of the LayerNorm and attention kernels inside one encoder forward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_rows_multi_kernel -s 60 -c 2 -o gpurun_out/prof_r2_ln \
  python tools/perf_encoder.py 592x128 > gpurun_out/h_ncu_ln.log 2>&1
echo "ncu ln rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 30 -c 2 -o gpurun_out/prof_r2_attn \
  python tools/perf_encoder.py 592x128 > gpurun_out/h_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 30 -c 1 -o gpurun_out/prof_r2_attn512 \
  python tools/perf_encoder.py 148x512 > gpurun_out/h_ncu_attn512.log 2>&1
echo "ncu attn512 rc=$?"
