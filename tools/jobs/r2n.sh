mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2n_pytest.log | cut -c1-300
timeout 700 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2n_bench_default.json; tail -c 400 gpurun_out/r2n_bench_default.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python bench.py --workload train --quick --steps 1 --warmup 1 > gpurun_out/r2n_plain_train.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_train256_launches.csv python bench.py --workload train --quick --steps 1 --warmup 1 > gpurun_out/r2n_ncu_train.log 2>&1; echo "ncu train rc=$?"
timeout 300 python bench.py --workload infer --quick --steps 1 --warmup 1 > gpurun_out/r2n_plain_infer.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_infer_launches.csv python bench.py --workload infer --quick --steps 1 --warmup 1 > gpurun_out/r2n_ncu_infer.log 2>&1; echo "ncu infer rc=$?"
timeout 200 python tools/ncu_targets.py > gpurun_out/r2n_plain_targets.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cm3p -o gpurun_out/r2_targets_full -f python tools/ncu_targets.py > gpurun_out/r2n_ncu_targets.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2n_ncu_targets.log
ls -la gpurun_out | head -30
