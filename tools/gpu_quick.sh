#!/usr/bin/env bash
# Quick device-resident timing of the short workload (+ optional GPU tests with PYTEST=1).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["kernel_ms"], d.get("records_by_kernel"))'
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
if [ "${PYTEST:-0}" = "1" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log; fi
echo "== default"; $B 2>&1 | python -c "$pick"
for v in ${VARIANTS:-}; do
  rm -f cactus-gfa-tools_b200/lib/libg2p.so
  make -s EXTRA_NVFLAGS="$v" cactus-gfa-tools_b200/lib/libg2p.so > /dev/null 2>&1
  echo "== $v"; $B 2>&1 | python -c "$pick"
done
if [ "${NCU:-0}" = "1" ]; then
  S="python bench.py --records 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:${NCU_K:-k_rec}" -s 3 -c 1 -f -o gpurun_out/prof_quick $S > gpurun_out/ncu_quick.log 2>&1
  echo "ncu rc=$?"
fi
