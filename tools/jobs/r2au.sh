mkdir -p gpurun_out
timeout 400 python bench.py --workload train --variations 256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_train_v256_1gpu_final.json 2> gpurun_out/r2au_v256.err; echo "v256 rc=$?"; python tools/print_bench.py gpurun_out/r2_bench_train_v256_1gpu_final.json; tail -c 200 gpurun_out/r2au_v256.err
timeout 300 python bench.py --workload mlm --steps 10 --warmup 3 > gpurun_out/r2_bench_mlm_1gpu_final.json 2> gpurun_out/r2au_mlm.err; echo "mlm rc=$?"; python tools/print_bench.py gpurun_out/r2_bench_mlm_1gpu_final.json
