mkdir -p gpurun_out
timeout 120 python tools/sanitize_cases.py > gpurun_out/sanitize_plain_racecheck.log 2>&1 && echo "plain run ok" && timeout 800 compute-sanitizer --tool racecheck --print-limit 40 python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "sanitizer rc=$?"; tail -25 gpurun_out/r2_sanitizer_racecheck.log | cut -c1-300
