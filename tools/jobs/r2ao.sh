mkdir -p gpurun_out
for f in 0 1 0 1; do echo "FUSE_LN=$f"; CM3P_FUSE_LN=$f timeout 300 python bench.py --workload infer --quick --steps 20 --warmup 5 2>/dev/null; done
CM3P_FUSE_LN=1 timeout 600 python -m pytest tests/test_model_parity_gpu.py tests/test_fullsize_properties_gpu.py -m gpu -q > gpurun_out/r2ao_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ao_pytest.log | cut -c1-200
