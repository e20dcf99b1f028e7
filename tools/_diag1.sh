set -x
cd $GRAFT_REPO_ROOT
python tools/census.py sdxl_unet_B2_1024 > gpurun_out/d1_census_sdxl.txt 2>&1
for shp in "2048 1280 1280" "8192 640 640" "528 3072 3072" "154 1280 2048" "2048 10240 1280"; do
  for r in 0 16; do
    R=$r python tools/tc_probe.py $shp >> gpurun_out/d1_probe.txt 2>&1
    R=$r VFT_TC_DEBUG=16 python tools/tc_probe.py $shp >> gpurun_out/d1_probe_tl.txt 2>&1
  done
done
