# 2-GPU box: full GPU suite, smoke, N=1 and N=2 bench with the ping-pong e2e leg, Tenengrad probe after software pipelining
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2j_smoke.log 2>&1
python scripts/sharp_probe.py > gpurun_out/r2j_sharp_probe.log 2>&1
STK_TENENGRAD_COLS=8 python scripts/sharp_probe.py > gpurun_out/r2j_sharp_probe_c8.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
timeout 300 $TR --nproc-per-node 2 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
tail -4 gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_smoke.log; head -4 gpurun_out/r2j_sharp_probe.log; head -2 gpurun_out/r2j_sharp_probe_c8.log; grep -h "^{" gpurun_out/r2j_bench_n1.json | cut -c1-200; tail -3 gpurun_out/r2j_bench_n1.err; grep -h "^{" gpurun_out/r2j_bench_n2.json | cut -c1-200; tail -3 gpurun_out/r2j_bench_n2.err
