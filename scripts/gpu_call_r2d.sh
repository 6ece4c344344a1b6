# K4 second generation: parity (warp tests + full-size config 5), A/B against generation 1, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "warp or config5 or config4 or keypoint or stack" > gpurun_out/r2d_pytest_warp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_warp.log
STK_WARP_GEN=1 python scripts/config5_scale.py --frames 64 --steps 3 --no-check > gpurun_out/r2d_cfg5_gen1.json 2> gpurun_out/r2d_cfg5_gen1.err
STK_WARP_GEN=2 python scripts/config5_scale.py --frames 64 --steps 3 > gpurun_out/r2d_cfg5_gen2.json 2> gpurun_out/r2d_cfg5_gen2.err
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
export STK_LOOP_MODE=host
CMD="python bench.py --frames 17 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2d_plain_host.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:warp_accumulate_v2 -s 4 -c 1 -o gpurun_out/prof_warp_r2d $CMD > gpurun_out/r2d_ncu_warp.log 2>&1
ncu -i gpurun_out/prof_warp_r2d.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_raw_warp.csv 2>/dev/null
ncu -i gpurun_out/prof_warp_r2d.ncu-rep --page source --csv --print-source sass > gpurun_out/r2d_ncu_src_warp.csv 2>/dev/null
unset STK_LOOP_MODE
tail -5 gpurun_out/r2d_pytest_warp.log; cat gpurun_out/r2d_cfg5_gen1.json gpurun_out/r2d_cfg5_gen2.json; tail -3 gpurun_out/r2d_cfg5_gen2.err; cat gpurun_out/r2d_bench.json | cut -c1-1500
