mkdir -p gpurun_out
CM3P_LIB_PATH=variants/libprof.so timeout 120 python tools/attn_one.py 64 bwd > gpurun_out/r2h_prof.log 2>&1; grep "win bwd" gpurun_out/r2h_prof.log | head -14
timeout 600 python -m pytest tests/test_kernels_bwd_gpu.py tests/test_round2_gpu.py -m gpu -q --maxfail=8 -k "attention or streaming or reproducible" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
timeout 200 python tools/bench_kernels.py winbwd > gpurun_out/r2h_winbwd_walk.jsonl 2>&1; cat gpurun_out/r2h_winbwd_walk.jsonl | grep attn_bwd
