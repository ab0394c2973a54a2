mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2c_pytest.log
timeout 300 python tools/train_probe.py --batches 256 --profile > gpurun_out/r2c_probe_det.log 2>&1
head -12 gpurun_out/r2c_probe_det.log | cut -c1-200
