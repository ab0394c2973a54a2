mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus 4 --steps 6 --warmup 3 --workload train --no-cpu-baseline 2> gpurun_out/r2as_train4.err | grep '^{' > gpurun_out/r2as_train_4gpu.json; echo "train rc=$?"
python tools/print_bench.py gpurun_out/r2as_train_4gpu.json; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2as_train_4gpu.json").read().strip().splitlines()[-1])
print(json.dumps({k:d["comm"][k] for k in ("all_reduce_ms","step_ms_without_collectives","step_ms_with_collectives","exposed_ms","exposed_tail_ms","overlap")}))
PY
tail -c 200 gpurun_out/r2as_train4.err
