mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_parity_gpu.py -m gpu -q -x > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ah_pytest.log | cut -c1-300
timeout 300 python tools/bench_kernels.py 2>&1 | grep gemm | cut -c1-150
timeout 300 python bench.py --workload infer --quick --steps 10 --warmup 3 2>/dev/null
timeout 300 python bench.py --workload train --quick --steps 5 --warmup 3 2>/dev/null
