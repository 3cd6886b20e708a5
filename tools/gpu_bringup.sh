#!/bin/bash
# First-contact run on the B200 box: every risky kernel family in its own process so that one
# trap / illegal access cannot mask the others.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv | tee gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
echo "=== non-conv kernels"; timeout 600 python -m pytest tests/test_gpu_kernels.py -k "not conv" -q -s -p no:cacheprovider 2>&1 | tee gpurun_out/k_other.log | tail -40
N=$(python - <<'PY'
import sys; sys.path.insert(0,'tests')
import test_gpu_kernels as t; print(len(t.CONV_CASES))
PY
)
echo "=== conv cases ($N), one process each"
: > gpurun_out/k_conv.log
for i in $(seq 0 $((N-1))); do
  timeout 180 python - "$i" >> gpurun_out/k_conv.log 2>&1 <<'PY'
import sys, traceback
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import torch
import test_gpu_kernels as t
from sgdm_b200 import _lib
i = int(sys.argv[1])
lib = _lib.lib(); lib._op = torch.float16 if lib.sgdm_operand_dtype() == b"f16" else torch.bfloat16
case = t.CONV_CASES[i]
try:
    t.test_conv_tcgen05_vs_torch(lib, case)
    print(f"CASE {i} PASS: {case[-1]}")
except BaseException as e:
    print(f"CASE {i} FAIL: {case[-1]}: {type(e).__name__}: {str(e)[:300]}")
PY
  rc=$?; [ $rc -ne 0 ] && echo "CASE $i process exit code $rc" >> gpurun_out/k_conv.log
done
grep -E "CASE|rel_l2" gpurun_out/k_conv.log | tail -80
echo "=== e2e"; timeout 1200 python -m pytest tests/test_gpu_e2e.py -q -s -p no:cacheprovider 2>&1 | tee gpurun_out/e2e.log | grep -E "rel_l2|PSNR|passed|failed|Error|error|FAILED|PASSED" | tail -80
