mkdir -p gpurun_out
timeout 120 python tools/sanitize_cases.py > gpurun_out/sanitize_plain_synccheck.log 2>&1 && echo "plain run ok" && timeout 1500 compute-sanitizer --tool synccheck --print-limit 40 python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_synccheck.log 2>&1; echo "sanitizer rc=$?"; tail -25 gpurun_out/r2_sanitizer_synccheck.log | cut -c1-300
