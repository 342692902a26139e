#!/bin/bash
# round-2 GPU call O (1 GPU): last sanity of the committed state: smoke() and the quick parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/o_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/o_smoke.log
timeout 600 python -m pytest tests -m gpu -q --timeout=500 -p no:cacheprovider -k "not full_size and not overlap and not tier2" > gpurun_out/o_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/o_pytest.log
