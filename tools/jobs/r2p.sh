mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_bwd_gpu.py tests/test_round2_gpu.py -m gpu -q -x -k "attention or attn or streaming or base_config or reproducible" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log | cut -c1-300
timeout 300 python tools/bench_kernels.py winbwd > gpurun_out/r2p_winbwd.jsonl 2>&1; cat gpurun_out/r2p_winbwd.jsonl | cut -c1-300
timeout 300 python bench.py --workload train --quick --steps 3 --warmup 2 > gpurun_out/r2p_train_quick.json 2> gpurun_out/r2p_train_quick.err; cat gpurun_out/r2p_train_quick.json; tail -c 300 gpurun_out/r2p_train_quick.err
