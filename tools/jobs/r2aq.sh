mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 700 python bench.py --steps 20 > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; echo "bench rc=$?"; python tools/print_bench.py gpurun_out/r2h_bench_default.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python bench.py --workload infer --quick --steps 1 --warmup 1 > gpurun_out/r2h_plain_infer.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_infer_launches_v4.csv python bench.py --workload infer --quick --steps 1 --warmup 1 > gpurun_out/r2h_ncu_infer.log 2>&1; echo "ncu infer rc=$?"
timeout 300 python bench.py --workload train --quick --steps 1 --warmup 1 > gpurun_out/r2h_plain_train.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_train256_launches_v4.csv python bench.py --workload train --quick --steps 1 --warmup 1 > gpurun_out/r2h_ncu_train.log 2>&1; echo "ncu train rc=$?"
