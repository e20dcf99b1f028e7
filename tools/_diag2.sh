cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_contraction or quantize_many or one_call_backward or layer_fwd_bwd" > gpurun_out/d2_pytest.txt 2>&1
tail -5 gpurun_out/d2_pytest.txt
timeout 300 python tools/quant_probe.py > gpurun_out/d2_quant.txt 2>&1; tail -4 gpurun_out/d2_quant.txt
for shp in "528 3072 3072" "264 3072 3072" "154 1280 2048" "2048 1280 1280" "16 3072 3072"; do
  for r in 0 16; do R=$r timeout 120 python tools/tc_probe.py $shp >> gpurun_out/d2_probe.txt 2>&1; done
done
cat gpurun_out/d2_probe.txt
timeout 300 python tools/step_probe.py > gpurun_out/d2_step.txt 2>&1; cat gpurun_out/d2_step.txt
timeout 600 python tools/census.py sdxl_unet_B2_1024 auraflow_6.8B_B2_1024 > gpurun_out/d2_census.txt 2>&1; cat gpurun_out/d2_census.txt
