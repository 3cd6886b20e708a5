#!/bin/bash
# Whole-step A/B on ONE box: runs bench.py once per environment setting given as arguments
# ("-" = defaults, otherwise "VAR=val[,VAR2=val2]"), interleaved ROUNDS times, and prints ms/step plus
# the per-family split of one profiled step.   tools/ab_step.sh - SGDM_PDL_CLASSES=1
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for round in $(seq 1 ${ROUNDS:-2}); do
for setting in "$@"; do
  tag=$(echo "$setting" | tr ',=' '__')
  envs=""
  [ "$setting" != "-" ] && envs=$(echo "$setting" | tr ',' ' ')
  env $envs timeout 600 python bench.py --config ${CONFIG:-2} --no-cpu-baseline --no-e2e --steps ${STEPS:-20} --warmup 3 --dump-ops gpurun_out/ops_$tag.json > gpurun_out/ab_$tag.log 2>&1
  python - "$setting" gpurun_out/ab_$tag.log <<'PY'
import json, sys
line = [l for l in open(sys.argv[2]) if l.startswith("{")]
if not line:
    print("==", sys.argv[1], "FAILED"); print(open(sys.argv[2]).read()[-1500:]); sys.exit(0)
d = json.loads(line[-1])
fam = {k: round(v["ms"], 3) for k, v in d["roofline"]["families"].items()}
print(f"== {sys.argv[1]:40s} {d['ms_per_step']:.3f} ms/step  clk {d['clocks']['sm_mhz']}  {fam}", flush=True)
PY
done
done
