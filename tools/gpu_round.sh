#!/bin/bash
# Standard GPU visit: the driver's test command, the bench line, an ncu launch list of one
# steady-state step and one full ncu capture of the dominant kernel.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -s 2>&1 | tee gpurun_out/pytest_gpu.log | grep -E "rel_l2|PSNR|passed|failed|FAILED|Error" | tail -60
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --dump-ops gpurun_out/ops.json ${BENCH_ARGS:-} 2>&1 | tee gpurun_out/bench.log | tail -3
if [ "${NCU:-1}" = "1" ]; then
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${NCU_SKIP:-600} --launch-count ${NCU_COUNT:-180} --csv --log-file gpurun_out/launches.csv python bench.py --ncu --steps 3 --warmup 3 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
echo "=== ncu full capture of the conv kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip ${NCU_CONV_SKIP:-260} --launch-count 4 -f -o gpurun_out/prof_conv python bench.py --ncu --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/
fi
