mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_round2_gpu.py tests/test_train_entry.py -m gpu -q --maxfail=5 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --workload train --train-batch 512 --variations 1 --quick > gpurun_out/r2k_b512.json 2> gpurun_out/r2k_b512.err; echo "b512 rc=$?"; cat gpurun_out/r2k_b512.json; tail -c 600 gpurun_out/r2k_b512.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
