mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest.log
timeout 400 python bench.py --steps 8 --warmup 3 --workload train --no-cpu-baseline 2> gpurun_out/r2j_train.err | grep '^{' > gpurun_out/r2j_train_v8.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_train_v8.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","optimizer_step_ms","peak_mem_gb","mfu")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["roofline_attention"]["achieved"], d["roofline_attention"]["share_of_step"], d["clocks"])
PY
