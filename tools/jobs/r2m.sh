mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_embed_tools_gpu.py -m gpu -q --maxfail=10 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2m_pytest.log | cut -c1-250
