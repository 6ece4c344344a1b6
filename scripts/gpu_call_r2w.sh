# 1 GPU: where does the Affine iteration kernel spend its time? (config 3's motion model)
mkdir -p gpurun_out
export STK_LOOP_MODE=host
CMD="python bench.py --motion 2 --frames 6 --steps 1 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2w_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ecc_iter_v2 -s 8 -c 1 -o gpurun_out/prof_ecc_affine_r2w $CMD > gpurun_out/r2w_ncu.log 2>&1
ncu -i gpurun_out/prof_ecc_affine_r2w.ncu-rep --page raw --csv > gpurun_out/r2w_ncu_raw_ecc_affine.csv 2>/dev/null
ncu -i gpurun_out/prof_ecc_affine_r2w.ncu-rep --page source --csv --print-source sass > gpurun_out/r2w_ncu_src_ecc_affine.csv 2>/dev/null
unset STK_LOOP_MODE
python bench.py --motion 2 --frames 32 --steps 3 --warmup 2 --skip-cpu --skip-e2e > gpurun_out/r2w_bench_affine.json 2> gpurun_out/r2w_bench_affine.err
cut -c1-200 gpurun_out/r2w_bench_affine.json; tail -2 gpurun_out/r2w_ncu.log
