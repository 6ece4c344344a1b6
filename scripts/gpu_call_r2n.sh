# 1 GPU: wider cross-block sum rounds — tail timing, one-lane figure, ECC parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "ecc or iteration or config or kernel_variants or lanes" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
python scripts/timing_probe.py 2:16 > gpurun_out/r2n_timing_probe.log 2>&1
python scripts/k2_variants.py 13 2:16 2:16:1 > gpurun_out/r2n_k2_variants.log 2>&1
tail -3 gpurun_out/r2n_pytest.log; grep -E "tail|variant|pixels" gpurun_out/r2n_timing_probe.log; cat gpurun_out/r2n_k2_variants.log
