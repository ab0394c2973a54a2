mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --workload train --no-cpu-baseline --grad-overlap on 2> gpurun_out/r2e_on.err | grep '^{' > gpurun_out/r2e_overlap_on.json; echo "on rc=$?"
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --workload train --no-cpu-baseline --grad-overlap off 2> gpurun_out/r2e_off.err | grep '^{' > gpurun_out/r2e_overlap_off.json; echo "off rc=$?"
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --workload train --no-cpu-baseline --grad-overlap on --nccl-max-ctas 4 2> gpurun_out/r2e_on4.err | grep '^{' > gpurun_out/r2e_overlap_on_ctas4.json; echo "on4 rc=$?"
python - <<'PY'
import json
for n in ("on","off","on_ctas4"):
    try:
        d=json.loads(open(f"gpurun_out/r2e_overlap_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], json.dumps({k:d["comm"][k] for k in ("all_reduce_ms","step_ms_without_collectives","step_ms_with_collectives","exposed_ms","overlap")}))
    except Exception as e: print(n, "ERR", e)
PY
tail -c 300 gpurun_out/r2e_on4.err
