mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.log | cut -c1-300
timeout 700 python bench.py --steps 20 > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; echo "bench rc=$?"; python tools/print_bench.py gpurun_out/r2g_bench_default.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2g_bench_reference.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python bench.py --workload train --quick --steps 1 --warmup 1 > gpurun_out/r2g_plain_train.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_train256_launches_v3.csv python bench.py --workload train --quick --steps 1 --warmup 1 > gpurun_out/r2g_ncu_train.log 2>&1; echo "ncu train rc=$?"
timeout 300 python bench.py --workload infer --quick --steps 1 --warmup 1 > gpurun_out/r2g_plain_infer.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_infer_launches_v3.csv python bench.py --workload infer --quick --steps 1 --warmup 1 > gpurun_out/r2g_ncu_infer.log 2>&1; echo "ncu infer rc=$?"
timeout 200 python tools/ncu_targets.py > gpurun_out/r2g_plain_targets.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:attn_|gemm_bf16|layernorm|geglu' -o gpurun_out/r2_targets_full_v3 -f python tools/ncu_targets.py > gpurun_out/r2g_ncu_targets.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2g_ncu_targets.log
