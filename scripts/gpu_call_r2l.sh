# 2-GPU box: exchange duration against the reduce kernel's grid (4K payload), then the N=2 bench with the default grid
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/r2l_peer_blocks.log
for B in 592 148 74 37 16; do
  echo "STK_PEER_BLOCKS=$B" >> gpurun_out/r2l_peer_blocks.log
  STK_PEER_BLOCKS=$B PC_W=3840 PC_H=2160 timeout 120 $TR --nproc-per-node 2 --master-port 2957$((B % 10)) scripts/peer_check.py 2>&1 | grep -E "per exchange|ok|Error|error" >> gpurun_out/r2l_peer_blocks.log
done
timeout 300 $TR --nproc-per-node 2 --master-port 29581 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err
STK_PEER_BLOCKS=592 timeout 300 $TR --nproc-per-node 2 --master-port 29582 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2l_bench_n2_b592.json 2> gpurun_out/r2l_bench_n2_b592.err
cat gpurun_out/r2l_peer_blocks.log; grep -h "^{" gpurun_out/r2l_bench_n2.json | cut -c1-200; grep -h "^{" gpurun_out/r2l_bench_n2_b592.json | cut -c1-200
