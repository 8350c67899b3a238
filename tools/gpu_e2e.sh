#!/usr/bin/env bash
# e2e (host buffers) timeline and chunk-size sweep of g2p_convert_host on the short workload.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6)'
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
echo "== default"; $B 2>gpurun_out/e2e_default.err | python -c "$pick"
for mb in ${CHUNKS:-32 64 192}; do echo "== chunk ${mb} MB"; G2P_HOST_CHUNK_MB=$mb $B 2>/dev/null | python -c "$pick"; done
echo "== trace"; G2P_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/e2e_trace.log; grep -c "g2p trace" gpurun_out/e2e_trace.log
