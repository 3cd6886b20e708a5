python -c "import __graft_entry__ as g; g.build()" | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "conv or stats" -q -p no:cacheprovider 2>&1 | tail -3
python tools/bench_conv.py "proj 512->512 1x1 @16 B512" "conv 128->128 3x3 @64 B512" "first 64->128 3x3 @64 B512" 2>&1 | tee gpurun_out/bench_conv.log | grep -E "stats=1|res=0 stats=0" 
NCU=0 bash tools/gpu_round.sh
