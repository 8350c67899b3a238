#!/usr/bin/env bash
# k_rec vs k_short on the short-read workload (device-resident timing), plus compile-time variants.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["kernel_ms"], d.get("records_by_kernel"))'
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
echo "== k_rec default"; $B 2>&1 | python -c "$pick"
echo "== k_short"; G2P_SIZE_KERNEL=short $B 2>&1 | python -c "$pick"
for c in 10 12 16; do echo "== k_rec chunks=$c"; G2P_REC_CHUNKS=$c $B 2>&1 | python -c "$pick"; done
for v in ${VARIANTS:-"-DG2P_REC_CTAS=3" "-DG2P_REC_CTAS=5"}; do
  rm -f cactus-gfa-tools_b200/lib/libg2p.so
  make -s EXTRA_NVFLAGS="$v" cactus-gfa-tools_b200/lib/libg2p.so > /dev/null 2>&1
  echo "== $v"; $B 2>&1 | python -c "$pick"
done
