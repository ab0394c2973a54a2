mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm" > gpurun_out/r2ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ar_pytest.log | cut -c1-300
timeout 200 python tools/stress_attn.py > gpurun_out/r2ar_stress.log 2>&1; echo "stress rc=$?"; tail -3 gpurun_out/r2ar_stress.log | cut -c1-200
