#!/bin/bash
# round-2 run M: the token-parallel kernels k_par_* on the GPU: the whole suite, then long-record workloads with and without them
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2m_pytest.log
tail -3 gpurun_out/r2m_pytest.log
Q="--no-cli --no-cpu-baseline --steps 6 --warmup 3"
for v in "G2P_PAR=1" "G2P_PAR=0"; do
  for w in stable medium asm mixed; do
    env $v timeout 400 python bench.py --workload $w $Q > gpurun_out/r2m_${w}_${v#G2P_PAR=}.json 2> gpurun_out/r2m_${w}_${v#G2P_PAR=}.err
    echo "$v $w rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2m_${w}_${v#G2P_PAR=}.json | head -1) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2m_${w}_${v#G2P_PAR=}.json | head -1) $(grep -o '"records_by_kernel": {[^}]*}' gpurun_out/r2m_${w}_${v#G2P_PAR=}.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2m_${w}_${v#G2P_PAR=}.json | head -1)"
  done
done
