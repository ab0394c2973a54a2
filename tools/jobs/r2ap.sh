mkdir -p gpurun_out
for f in 0 1 0 1; do echo "FUSE_LN=$f"; CM3P_FUSE_LN=$f timeout 300 python bench.py --workload train --quick --steps 6 --warmup 3 2>/dev/null; done
CM3P_FUSE_LN=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2ap_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ap_pytest.log | cut -c1-200
