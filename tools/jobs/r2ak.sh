mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_kernels_bwd_gpu.py -m gpu -q -x -k "gemm or conv or linear or wgrad or split" > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ak_pytest.log | cut -c1-300
timeout 300 python tools/bench_kernels.py 2>&1 | grep gemm | cut -c1-150
timeout 300 python tools/bench_kernels.py gemmbwd 2>&1 | grep kernel | cut -c1-150
