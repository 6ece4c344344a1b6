# 1 GPU: world-of-one exchange (fused lane sum + divide on the exchange stream), steps pipelined at N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py tests/test_gpu_parity.py -m gpu -q -k "peer or world or back_to_back or lanes or cache or stream" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
STK_REDUCE=nccl timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2t_bench_nopipe.json 2> gpurun_out/r2t_bench_nopipe.err
tail -3 gpurun_out/r2t_pytest.log; cut -c1-220 gpurun_out/r2t_bench.json; tail -2 gpurun_out/r2t_bench.err; cut -c1-220 gpurun_out/r2t_bench_nopipe.json
