mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_muon_gpu.py -q -x > gpurun_out/r2b_muon.log 2>&1; tail -3 gpurun_out/r2b_muon.log
timeout 300 python tools/train_probe.py --batches 256 --profile > gpurun_out/r2b_probe_det.log 2>&1
timeout 300 python tools/train_probe.py --batches 256 --profile --opt 4=0 > gpurun_out/r2b_probe_atomic.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --workload train --variations 256 --no-cpu-baseline > gpurun_out/r2b_train_v256.json 2> gpurun_out/r2b_train_v256.err; echo "v256 rc=$?"; tail -c 1500 gpurun_out/r2b_train_v256.err
timeout 500 python tools/bench_reference_gpu.py --steps 3 > gpurun_out/r2b_reference_gpu.jsonl 2> gpurun_out/r2b_reference_gpu.err; cat gpurun_out/r2b_reference_gpu.jsonl
head -30 gpurun_out/r2b_probe_det.log
