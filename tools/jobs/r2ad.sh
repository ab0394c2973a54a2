mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ad_pytest.log | cut -c1-300
timeout 700 python bench.py > gpurun_out/r2ad_bench_default.json 2> gpurun_out/r2ad_bench_default.err; echo "bench rc=$?"; python tools/print_bench.py gpurun_out/r2ad_bench_default.json; tail -c 300 gpurun_out/r2ad_bench_default.err
