# round 2, second session, first call: everything that was measured in the lost session, once more
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_gpus.log
python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
scripts/pipe_probe > gpurun_out/r2b_pipe_probe.log 2>&1
python scripts/k2_variants.py 13 1:0 2:0 2:2 2:2:1 2:8 2:9 2:12 2:13 2:1 2:10 2:11 2:4 > gpurun_out/r2b_k2_variants.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python scripts/sharp_probe.py > gpurun_out/r2b_sharp_probe.log 2>&1
export STK_LOOP_MODE=host
CMD="python bench.py --frames 8 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2b_plain_host.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b_launches_host.csv $CMD > gpurun_out/r2b_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ecc_iter_v2 -s 12 -c 2 -o gpurun_out/prof_ecc_r2b $CMD > gpurun_out/r2b_ncu_ecc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_accumulate -s 2 -c 1 -o gpurun_out/prof_warp_r2b $CMD > gpurun_out/r2b_ncu_warp.log 2>&1
unset STK_LOOP_MODE
tail -3 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_pipe_probe.log gpurun_out/r2b_k2_variants.log; tail -2 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json
