mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x > gpurun_out/r2l_dist8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_dist8_pytest.log
timeout 500 $TR --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 --workload train --no-cpu-baseline 2> gpurun_out/r2l_train8.err | grep '^{' > gpurun_out/r2l_train_8gpu.json; echo "train rc=$?"
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --workload train --global-negatives --train-batch 512 --variations 1 --no-cpu-baseline 2> gpurun_out/r2l_glob8.err | grep '^{' > gpurun_out/r2l_globalneg_8gpu.json; echo "globalneg rc=$?"
timeout 400 $TR --master-port 29523 bench.py --gpus 8 --steps 8 --warmup 3 --workload mlm 2> gpurun_out/r2l_mlm8.err | grep '^{' > gpurun_out/r2l_mlm_8gpu.json; echo "mlm rc=$?"
timeout 500 $TR --master-port 29524 bench.py --gpus 8 --steps 5 --warmup 3 --workload train --variations 256 --no-cpu-baseline 2> gpurun_out/r2l_v256.err | grep '^{' > gpurun_out/r2l_train_v256_8gpu.json; echo "v256 rc=$?"
python - <<'PY'
import json
for n in ("train_8gpu","globalneg_8gpu","mlm_8gpu","train_v256_8gpu"):
    try:
        d=json.loads(open(f"gpurun_out/r2l_{n}.json").read().strip().splitlines()[-1])
        c=d.get("comm") or {}
        print(n, d["value"], d["unit"], d["ms_per_step"], "e2e", d["e2e"]["value"], "mem", d["peak_mem_gb"], "opt", d.get("optimizer_step_ms"), "ar", c.get("all_reduce_ms"), "exposed", c.get("exposed_ms"), "overlap", c.get("overlap"))
    except Exception as e: print(n, "ERR", e)
PY
tail -c 300 gpurun_out/r2l_glob8.err
