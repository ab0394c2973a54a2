mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x > gpurun_out/r2d_dist2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2d_dist2_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --workload train > gpurun_out/r2d_train_2gpu.json 2> gpurun_out/r2d_train_2gpu.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2d_train_2gpu.err; cut -c1-600 gpurun_out/r2d_train_2gpu.json
