mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_bwd_gpu.py tests/test_round2_gpu.py -m gpu -q -x -k "attention or attn or streaming or base_config or reproducible" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z_pytest.log | cut -c1-300
CM3P_LIB_PATH=variants/libprof.so timeout 120 python tools/attn_one.py 64 bwd > gpurun_out/r2z_prof.log 2>&1; echo "rc=$?"; grep "win bwd" gpurun_out/r2z_prof.log | sort | tail -24 | cut -c1-420
timeout 300 python tools/bench_kernels.py winbwd > gpurun_out/r2z_winbwd.jsonl 2>&1; grep attn_bwd gpurun_out/r2z_winbwd.jsonl | cut -c1-300
