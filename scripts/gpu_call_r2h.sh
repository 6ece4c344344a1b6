# 2-GPU box: full GPU suite after the device-side ordering rework (non-blocking reset / set_reference, exchange stream, input stream), benches
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
timeout 300 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
timeout 300 $TR --nproc-per-node 2 --master-port 29542 scripts/config5_scale.py --frames 64 --steps 3 > gpurun_out/r2h_cfg5_n2.json 2> gpurun_out/r2h_cfg5_n2.err
python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1
tail -4 gpurun_out/r2h_pytest.log; grep -h "^{" gpurun_out/r2h_bench_n1.json | cut -c1-300; grep -h "^{" gpurun_out/r2h_bench_n2.json | cut -c1-300; tail -3 gpurun_out/r2h_bench_n2.err; grep -h "^{" gpurun_out/r2h_cfg5_n2.json; tail -2 gpurun_out/r2h_smoke.log
