#!/bin/bash
# ncu visit: (1) launch list of one steady-state step, (2) full capture of the conv_gemm launches of one
# step, (3) full capture of some gn_apply launches.  Only CSV exports travel back (reports are > 64 MiB).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
SKIP=${NCU_SKIP:-700}
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP --launch-count 175 --csv --log-file gpurun_out/launches.csv python bench.py --ncu --steps 3 --warmup 3 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
echo "=== ncu full: conv_gemm launches of one step"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip ${NCU_CONV_SKIP:-228} --launch-count ${NCU_CONV_COUNT:-57} -f -o /tmp/prof_conv python bench.py --ncu --steps 3 --warmup 3 > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/conv_raw.csv 2>/dev/null
for id in ${NCU_SRC_IDS:-3 4 11 41}; do
  ncu -i /tmp/prof_conv.ncu-rep --page source --csv --kernel-id :::$((id+1)) > gpurun_out/conv_src_$id.csv 2>/dev/null
done
if [ "${NCU_GN:-1}" = "1" ]; then
echo "=== ncu full: gn_apply launches"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_apply --launch-skip 196 --launch-count 12 -f -o /tmp/prof_gn python bench.py --ncu --steps 3 --warmup 3 > gpurun_out/ncu_gn.log 2>&1
tail -2 gpurun_out/ncu_gn.log
ncu -i /tmp/prof_gn.ncu-rep --page raw --csv > gpurun_out/gn_raw.csv 2>/dev/null
ncu -i /tmp/prof_gn.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/gn_src_1.csv 2>/dev/null
fi
ls -la gpurun_out/ /tmp/*.ncu-rep
du -sh gpurun_out
