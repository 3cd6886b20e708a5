#!/bin/bash
# Round-2 ncu visit on the final build (one GPU, never multi-rank):
#  (1) launch list (gpu__time_duration.sum, --clock-control none) of the WHOLE `bench.py --ncu --steps 3 --warmup 3`
#      run; tools/ncu_launch_list.py cuts out the last step and prints the per-family shares
#  (2) `--set full` of every conv_gemm launch of the last step (57), every gn_apply launch (49), the tcgen05 attention
#      launches (6); raw pages exported as CSV and condensed by tools/ncu_summary.py (reports are > 64 MiB)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
RUN="python bench.py --ncu --steps 3 --warmup 3"
STEPS=6   # 3 warm-up + 3 timed
echo "=== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all.csv $RUN > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_all.csv
python tools/ncu_launch_list.py gpurun_out/launches_all.csv gpurun_out/r02_ncu_launch_list.csv
echo "=== ncu full: conv_gemm launches of the last step"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip $((57 * (STEPS - 1))) --launch-count 57 -f -o /tmp/prof_conv $RUN > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/conv_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/conv_raw.csv gpurun_out/r02_ncu_conv_gemm.csv
for id in ${NCU_SRC_IDS:-11 17}; do   # the two Cout=128 residual convs at 64x64 (launch index within the step)
  ncu -i /tmp/prof_conv.ncu-rep --page source --csv --kernel-id :::$((id+1)) > gpurun_out/conv_src_$id.csv 2>/dev/null
  python tools/ncu_src_top.py gpurun_out/conv_src_$id.csv 20 > gpurun_out/r02_ncu_src_conv_$id.txt 2>/dev/null
done
echo "=== ncu full: gn_apply launches of the last step"
timeout 1200 ncu --set full --clock-control none -k regex:gn_apply --launch-skip $((49 * (STEPS - 1))) --launch-count 49 -f -o /tmp/prof_gn $RUN > gpurun_out/ncu_gn.log 2>&1
tail -1 gpurun_out/ncu_gn.log
ncu -i /tmp/prof_gn.ncu-rep --page raw --csv > gpurun_out/gn_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/gn_raw.csv gpurun_out/r02_ncu_gn_apply.csv
echo "=== ncu full: attn_tc2 launches of the last step"
timeout 900 ncu --set full --clock-control none -k regex:attn_tc2 --launch-skip $((6 * (STEPS - 1))) --launch-count 6 -f -o /tmp/prof_attn $RUN > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/attn_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/attn_raw.csv gpurun_out/r02_ncu_attn_tc2.csv
rm -f gpurun_out/conv_raw.csv gpurun_out/gn_raw.csv gpurun_out/attn_raw.csv gpurun_out/conv_src_*.csv
ls -la gpurun_out/r02_ncu_* ; du -sh gpurun_out
