mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_kernels_bwd_gpu.py tests/test_round2_gpu.py -m gpu -q -x -k "attention or attn or streaming or base_config or reproducible" > gpurun_out/r2an_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2an_pytest.log | cut -c1-300
CM3P_BENCH_B=256 timeout 300 python tools/bench_kernels.py attn bwd 2>&1 | grep '"attn_' | cut -c1-150
timeout 300 python tools/bench_kernels.py attn 2>&1 | grep '"attn_fwd' | cut -c1-150
