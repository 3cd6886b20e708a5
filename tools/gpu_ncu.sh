#!/bin/bash
# ncu visit: (1) launch list of one steady-state step, (2) full capture of every conv_gemm launch of one
# step, (3) full capture of the gn_apply launches of one step.  Reports land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
SKIP=${NCU_SKIP:-700}
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP --launch-count 175 --csv --log-file gpurun_out/launches.csv python bench.py --ncu --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
echo "=== ncu full: conv_gemm launches of one step"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip ${NCU_CONV_SKIP:-228} --launch-count ${NCU_CONV_COUNT:-57} -f -o gpurun_out/prof_conv python bench.py --ncu --steps 2 --warmup 3 > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
echo "=== ncu full: gn_apply launches of one step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_apply --launch-skip 196 --launch-count 49 -f -o gpurun_out/prof_gn python bench.py --ncu --steps 2 --warmup 3 > gpurun_out/ncu_gn.log 2>&1
tail -2 gpurun_out/ncu_gn.log
ls -la gpurun_out/
