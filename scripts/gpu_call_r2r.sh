# 2-GPU box, final state: full suite (peer tests included) and the N=2 bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
timeout 300 $TR --nproc-per-node 2 --master-port 29591 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err
tail -3 gpurun_out/r2r_pytest.log; grep -h "^{" gpurun_out/r2r_bench_n2.json | cut -c1-200; tail -2 gpurun_out/r2r_bench_n2.err
