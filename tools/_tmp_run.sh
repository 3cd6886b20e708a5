python -c "import __graft_entry__ as g; g.build()" | tail -1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider 2>&1 | tail -2
for cfg in "SGDM_ATTN_TC=1" "SGDM_ATTN_TC=0" "SGDM_ATTN_TC=1"; do
  echo "== $cfg"; env $cfg python bench.py --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], {k:v['ms'] for k,v in d['roofline']['families'].items()})"
done
