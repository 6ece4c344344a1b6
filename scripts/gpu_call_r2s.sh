# 8-GPU box, final state: N=8 bench (value + e2e)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err
grep -h "^{" gpurun_out/r2s_bench_n8.json | cut -c1-260; tail -2 gpurun_out/r2s_bench_n8.err
