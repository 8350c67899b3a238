#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { env "$@" timeout 300 python bench.py $Q > gpurun_out/r2j.out 2> gpurun_out/r2j.err; echo "$* $W rc $? $(grep -o 'unspecified launch failure\|illegal memory access\|gave up a spin wait.*' gpurun_out/r2j.err | sort | uniq -c | head -3) $(grep -o '\"ms_per_step\": [0-9.]*' gpurun_out/r2j.out | head -1)"; }
for W in medium stable asm; do
  Q="--no-cli --no-cpu-baseline --steps 8 --warmup 3 --workload $W"
  run G2P_FUSE=1; run G2P_FUSE=1; run G2P_FUSE=1 CUDA_LAUNCH_BLOCKING=1; run G2P_FUSE=2
done
W=short; Q="--no-cli --no-cpu-baseline --steps 10 --warmup 3 --workload short"
run G2P_FUSE=1; run G2P_FUSE=1 G2P_REC_WALK=nested
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2j.out").read().strip().splitlines()[-1]); print("nested", d["kernel_ms"])
P
G2P_FUSE=1 timeout 300 python bench.py $Q > gpurun_out/r2j_short.json 2>/dev/null; python - <<'P'
import json
d=json.loads(open("gpurun_out/r2j_short.json").read().strip().splitlines()[-1]); print("converged", d["kernel_ms"], d["e2e"]["ms_per_step"])
P
