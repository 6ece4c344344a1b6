# 1 GPU, final state: full suite, smoke, full bench (CPU baseline included), reference arm, ncu launch list of the host-loop bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err
export STK_LOOP_MODE=host
CMD="python bench.py --frames 8 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2o_plain_host.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2o_launches_host.csv $CMD > gpurun_out/r2o_ncu_list.log 2>&1
unset STK_LOOP_MODE
tail -4 gpurun_out/r2o_pytest.log; tail -3 gpurun_out/r2o_smoke.log; cut -c1-300 gpurun_out/r2o_bench.json; tail -2 gpurun_out/r2o_bench.err; cut -c1-300 gpurun_out/r2o_bench_ref.json
