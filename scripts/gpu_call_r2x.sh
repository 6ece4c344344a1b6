# 1 GPU: packed Affine accumulator — parity tests on every ECC path, config bench, affine bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "ecc or iteration or config or kernel_variants or lanes or scaling" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
python scripts/config_bench.py > gpurun_out/r2x_config_bench.log 2>&1
python bench.py --motion 2 --frames 32 --steps 3 --warmup 2 --skip-cpu --skip-e2e > gpurun_out/r2x_bench_affine.json 2> gpurun_out/r2x_bench_affine.err
tail -4 gpurun_out/r2x_pytest.log; cat gpurun_out/r2x_config_bench.log; cut -c1-200 gpurun_out/r2x_bench_affine.json
