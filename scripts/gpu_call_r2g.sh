# 8-GPU box: multi-GPU tests on real hardware, N=8 / N=4 bench with the per-rank diagnostic, configs[4] at scale, 4K exchange at world 8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 python -m pytest tests/test_gpu_peer_exchange.py -m gpu -q -x > gpurun_out/r2g_pytest_peer.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_peer.log
timeout 200 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err
STK_BENCH_TRACE=1 timeout 200 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 2 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2g_bench_n8_trace.json 2> gpurun_out/r2g_bench_n8_trace.err
timeout 200 $TR --nproc-per-node 4 --master-port 29533 bench.py --gpus 4 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2g_bench_n4.json 2> gpurun_out/r2g_bench_n4.err
timeout 200 $TR --nproc-per-node 2 --master-port 29536 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
timeout 240 $TR --nproc-per-node 8 --master-port 29534 scripts/config5_scale.py --frames 256 --steps 3 > gpurun_out/r2g_cfg5_n8.json 2> gpurun_out/r2g_cfg5_n8.err
PC_W=3840 PC_H=2160 timeout 120 $TR --nproc-per-node 8 --master-port 29535 scripts/peer_check.py > gpurun_out/r2g_peer_check_4k_world8.log 2>&1
tail -3 gpurun_out/r2g_pytest_peer.log; cat gpurun_out/r2g_bench_n8.json | cut -c1-700; tail -2 gpurun_out/r2g_bench_n8.err; cat gpurun_out/r2g_cfg5_n8.json; tail -5 gpurun_out/r2g_peer_check_4k_world8.log
