# K4 v2c (phased loads)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "warp or config5 or config4 or keypoint or stack" > gpurun_out/r2f_pytest_warp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_warp.log
python scripts/config5_scale.py --frames 64 --steps 3 > gpurun_out/r2f_cfg5_gen2.json 2> gpurun_out/r2f_cfg5_gen2.err
python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
export STK_LOOP_MODE=host
CMD="python bench.py --frames 17 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2f_plain_host.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:warp_accumulate_v2 -s 4 -c 1 -o gpurun_out/prof_warp_r2f $CMD > gpurun_out/r2f_ncu_warp.log 2>&1
ncu -i gpurun_out/prof_warp_r2f.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_raw_warp.csv 2>/dev/null
ncu -i gpurun_out/prof_warp_r2f.ncu-rep --page source --csv --print-source sass > gpurun_out/r2f_ncu_src_warp.csv 2>/dev/null
unset STK_LOOP_MODE
tail -5 gpurun_out/r2f_pytest_warp.log; cat gpurun_out/r2f_cfg5_gen2.json; cat gpurun_out/r2f_bench.json | cut -c1-600
