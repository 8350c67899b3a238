#!/usr/bin/env bash
# Rebuild libg2p.so on the GPU box with different compile-time settings and time the short workload.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in ${VARIANTS:-"-DG2P_SHORT_CTAS=4" "-DG2P_SHORT_CTAS=5" "-DG2P_SHORT_CTAS=6"}; do
  rm -f cactus-gfa-tools_b200/lib/libg2p.so
  make -s EXTRA_NVFLAGS="$v" cactus-gfa-tools_b200/lib/libg2p.so > /dev/null 2>&1
  echo "== $v"
  python bench.py ${BENCH_ARGS:---records 4000000 --steps 5} --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['kernel_ms'])
"
done
