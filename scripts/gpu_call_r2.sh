python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
python scripts/sharp_probe.py > gpurun_out/r2_sharp_probe2.log 2>&1
STK_TENENGRAD_COLS=8 python scripts/sharp_probe.py > gpurun_out/r2_sharp_probe2_c8.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
export STK_LOOP_MODE=host
CMD="python bench.py --frames 8 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2_plain_host2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2_host.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
$CMD > gpurun_out/r2_plain_host3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ecc_iter_v2 -s 12 -c 2 -o gpurun_out/prof_ecc_r2b $CMD > gpurun_out/r2_ncu_d.log 2>&1
$CMD > gpurun_out/r2_plain_host4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:prep_stream -s 2 -c 1 -o gpurun_out/prof_prep_r2 $CMD > gpurun_out/r2_ncu_e.log 2>&1
unset STK_LOOP_MODE
CMD2="python scripts/sharp_probe.py"
$CMD2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tenengrad_stream -s 2 -c 1 -o gpurun_out/prof_teng_r2 $CMD2 > gpurun_out/r2_ncu_f.log 2>&1
tail -3 gpurun_out/r2_pytest6.log; cat gpurun_out/r2_sharp_probe2.log gpurun_out/r2_sharp_probe2_c8.log; tail -2 gpurun_out/r2_bench_d.err
