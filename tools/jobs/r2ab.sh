mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_kernels_bwd_gpu.py -m gpu -q -x -k "gemm or conv or linear or wgrad or split" > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2ab_pytest.log | cut -c1-300
timeout 300 python tools/bench_kernels.py > gpurun_out/r2ab_kernels_mma2.jsonl 2>&1; grep gemm gpurun_out/r2ab_kernels_mma2.jsonl | cut -c1-200
timeout 300 python tools/bench_kernels.py --opt=2=3 > gpurun_out/r2ab_kernels_mcast.jsonl 2>&1; grep gemm gpurun_out/r2ab_kernels_mcast.jsonl | cut -c1-200
timeout 300 python tools/bench_kernels.py gemmbwd > gpurun_out/r2ab_gemmbwd_mma2.jsonl 2>&1; cat gpurun_out/r2ab_gemmbwd_mma2.jsonl | cut -c1-200
