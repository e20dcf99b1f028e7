cd $GRAFT_REPO_ROOT
rm -f gpurun_out/d3_*
for shp in "528 3072 3072" "264 3072 3072" "154 1280 2048" "154 640 2048" "77 1280 2048" "16 3072 3072" "2 18432 3072" "512 3072 2048"; do
  R=0 timeout 120 python tools/tc_probe.py $shp >> gpurun_out/d3_probe.txt 2>&1
  R=0 VFT_TC2_NOSPLIT=1 timeout 120 python tools/tc_probe.py $shp >> gpurun_out/d3_probe_nosplit.txt 2>&1
done
cat gpurun_out/d3_probe.txt; echo; cat gpurun_out/d3_probe_nosplit.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_contraction or layer_fwd_bwd" 2>&1 | tail -3
