cd $GRAFT_REPO_ROOT
set -x
python tools/job_once.py > gpurun_out/r02_job_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:qlora_tc2 -o gpurun_out/r02_prof_tc2_fused -f python tools/job_once.py > gpurun_out/r02_ncu_tc2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_tc2_fused.ncu-rep "qlora_tc2_kernel with the adapter inside: forward <side product>, backward <side product + dA/dB job>, config #1 (T=4096, 3072x3072, r=16)" > gpurun_out/r02_ncu_tc2_fused.txt 2>&1
python tools/quant_many_once.py > gpurun_out/r02_qm_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:nf4_quantize64 -c 2 -o gpurun_out/r02_prof_quant_many -f python tools/quant_many_once.py > gpurun_out/r02_ncu_qm.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_quant_many.ncu-rep "nf4_quantize64_kernel, one grouped launch: 96 x [3072, 3072] fp16 (vft_nf4_quantize_many)" > gpurun_out/r02_ncu_quant_many.txt 2>&1
python bench.py --steps 8 --warmup 3 --no-census --no-aura-step --no-cpu-baseline > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 8 --warmup 3 --no-census --no-aura-step --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
tail -3 gpurun_out/r02_ncu_tc2_fused.txt gpurun_out/r02_ncu_quant_many.txt
# small-launch evidence: SDXL C1280 attention projection (one wave, fixed-cost bound) and the split contraction (T = 16)
python tools/tc_once.py 2048 1280 1280 1 > gpurun_out/r02_small_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:qlora_tc2 -o gpurun_out/r02_prof_tc2_small -f python tools/tc_once.py 2048 1280 1280 1 > gpurun_out/r02_ncu_small.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_tc2_small.ncu-rep "qlora_tc2_kernel, NF4-only forward / backward at SDXL C1280 (T=2048, 1280x1280): one wave, fixed-cost bound" > gpurun_out/r02_ncu_tc2_small.txt 2>&1
python tools/tc_once.py 16 3072 3072 1 > gpurun_out/r02_split_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:qlora_tc2 -o gpurun_out/r02_prof_tc2_split -f python tools/tc_once.py 16 3072 3072 1 > gpurun_out/r02_ncu_split.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_tc2_split.ncu-rep "split contraction at T=16, 3072x3072: qlora_tc2_kernel (72 work items, fp32 slices) + qlora_tc2_finalize_kernel, forward / backward" > gpurun_out/r02_ncu_tc2_split.txt 2>&1
