#!/usr/bin/env bash
# File -> file timing of the drop-in gaf2paf executable on the GPU box next to the reference binary.
set -u
R="${GRAFT_REPO_ROOT:-/root/repo}"
cd /dev/shm
"$R/build/gafgen" short 10000000 a.gaf l.tsv 2>&1 | tail -1
E="$R/cactus-gfa-tools_b200/bin"
t() { local s=$(date +%s.%N); "$@"; local e=$(date +%s.%N); echo "wall $(echo "$e - $s" | bc -l | cut -c1-6) s"; }
echo "== gaf2paf (B200) 10M records -> /dev/null"; for i in 1 2; do t env G2P_STATS=1 "$E/gaf2paf" -l l.tsv a.gaf > /dev/null; done
echo "== gaf2paf (B200) 10M records -> tmpfs file"; t sh -c "$E/gaf2paf -l l.tsv a.gaf > out.paf"; ls -la out.paf
head -c 200000000 a.gaf | head -n -1 > s.gaf
echo "== reference gaf2paf, $(wc -l < s.gaf) records -> tmpfs file"; t sh -c "$R/oracle/_ref/gaf2paf -l l.tsv s.gaf > ref.paf"
"$E/gaf2paf" -l l.tsv s.gaf | cmp - ref.paf && echo CLI_PARITY_OK
rm -f a.gaf out.paf s.gaf ref.paf l.tsv
