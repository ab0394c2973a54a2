mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_bwd_gpu.py tests/test_round2_gpu.py -m gpu -q --maxfail=8 -k "attention or streaming or reproducible" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2f_pytest.log
timeout 200 python tools/bench_kernels.py winbwd > gpurun_out/r2f_winbwd_walk.jsonl 2>&1; cat gpurun_out/r2f_winbwd_walk.jsonl | grep attn_bwd
timeout 200 python tools/bench_kernels.py winbwd --opt=6=0 > gpurun_out/r2f_winbwd_v3.jsonl 2>&1; cat gpurun_out/r2f_winbwd_v3.jsonl | grep attn_bwd
