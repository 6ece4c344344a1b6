# 1 GPU: full GPU suite + smoke + bench after the asynchronous reset/set_reference rework; Tenengrad ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2i_smoke.log 2>&1
python scripts/sharp_probe.py > gpurun_out/r2i_sharp_probe.log 2>&1
CMD2="python scripts/sharp_probe.py"
ncu --set full --clock-control none --import-source on -k regex:tenengrad_stream -s 2 -c 1 -o gpurun_out/prof_teng_r2i $CMD2 > gpurun_out/r2i_ncu_teng.log 2>&1
ncu -i gpurun_out/prof_teng_r2i.ncu-rep --page raw --csv > gpurun_out/r2i_ncu_raw_teng.csv 2>/dev/null
ncu -i gpurun_out/prof_teng_r2i.ncu-rep --page source --csv --print-source sass > gpurun_out/r2i_ncu_src_teng.csv 2>/dev/null
export STK_LOOP_MODE=host
CMD="python bench.py --frames 8 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
ncu --set full --clock-control none --import-source on -k regex:prep_stream -s 3 -c 1 -o gpurun_out/prof_prep_r2i $CMD > gpurun_out/r2i_ncu_prep.log 2>&1
ncu -i gpurun_out/prof_prep_r2i.ncu-rep --page raw --csv > gpurun_out/r2i_ncu_raw_prep.csv 2>/dev/null
ncu -i gpurun_out/prof_prep_r2i.ncu-rep --page source --csv --print-source sass > gpurun_out/r2i_ncu_src_prep.csv 2>/dev/null
unset STK_LOOP_MODE
tail -4 gpurun_out/r2i_pytest.log; tail -3 gpurun_out/r2i_smoke.log; cat gpurun_out/r2i_sharp_probe.log
