# 8-GPU box after the device-side ordering rework: N=8 / N=4 bench (value + ping-pong e2e), lanes variant, configs[4] at scale
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err
timeout 240 $TR --nproc-per-node 4 --master-port 29562 bench.py --gpus 4 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err
timeout 200 $TR --nproc-per-node 8 --master-port 29563 bench.py --gpus 8 --steps 5 --warmup 3 --skip-cpu --skip-e2e --lanes 8 > gpurun_out/r2k_bench_n8_lanes8.json 2> gpurun_out/r2k_bench_n8_lanes8.err
timeout 240 $TR --nproc-per-node 8 --master-port 29564 scripts/config5_scale.py --frames 256 --steps 3 > gpurun_out/r2k_cfg5_n8.json 2> gpurun_out/r2k_cfg5_n8.err
grep -h "^{" gpurun_out/r2k_bench_n8.json | cut -c1-260; tail -2 gpurun_out/r2k_bench_n8.err; grep -h "^{" gpurun_out/r2k_bench_n4.json | cut -c1-200; grep -h "^{" gpurun_out/r2k_bench_n8_lanes8.json | cut -c1-200; grep -h "^{" gpurun_out/r2k_cfg5_n8.json
