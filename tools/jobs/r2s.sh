mkdir -p gpurun_out
timeout 120 python tools/attn_one.py 64 bwd > gpurun_out/r2s_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_win_kernel -o gpurun_out/r2s_win -f python tools/attn_one.py 64 bwd > gpurun_out/r2s_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s_ncu.log
