# K4 v2 (block-shared constants) + 512-thread K2 geometries
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "warp or config5 or config4 or keypoint or stack" > gpurun_out/r2e_pytest_warp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_warp.log
python scripts/config5_scale.py --frames 64 --steps 3 > gpurun_out/r2e_cfg5_gen2.json 2> gpurun_out/r2e_cfg5_gen2.err
python scripts/k2_variants.py 13 2:16 2:18 2:19 > gpurun_out/r2e_k2_variants.log 2>&1
python scripts/timing_probe.py 2:18 2:19 > gpurun_out/r2e_timing_probe.log 2>&1
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
export STK_LOOP_MODE=host
CMD="python bench.py --frames 17 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2e_plain_host.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:warp_accumulate_v2 -s 4 -c 1 -o gpurun_out/prof_warp_r2e $CMD > gpurun_out/r2e_ncu_warp.log 2>&1
ncu -i gpurun_out/prof_warp_r2e.ncu-rep --page raw --csv > gpurun_out/r2e_ncu_raw_warp.csv 2>/dev/null
ncu -i gpurun_out/prof_warp_r2e.ncu-rep --page source --csv --print-source sass > gpurun_out/r2e_ncu_src_warp.csv 2>/dev/null
unset STK_LOOP_MODE
tail -5 gpurun_out/r2e_pytest_warp.log; cat gpurun_out/r2e_cfg5_gen2.json; cat gpurun_out/r2e_k2_variants.log gpurun_out/r2e_timing_probe.log | cut -c1-600; cat gpurun_out/r2e_bench.json | cut -c1-900
