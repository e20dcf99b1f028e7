cd $GRAFT_REPO_ROOT
timeout 300 python tools/quant_probe.py 2>&1 | grep -E "AuraFlow|18432|3072, 3072" > gpurun_out/d4_q4.txt
VFT_QUANT_OCC=3 timeout 300 python tools/quant_probe.py 2>&1 | grep -E "AuraFlow|18432|3072, 3072" > gpurun_out/d4_q3.txt
cat gpurun_out/d4_q4.txt; echo; cat gpurun_out/d4_q3.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantize" 2>&1 | tail -3
