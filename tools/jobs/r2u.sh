mkdir -p gpurun_out
CM3P_LIB_PATH=variants/libprof.so timeout 120 python tools/attn_one.py 64 bwd > gpurun_out/r2u_prof.log 2>&1; echo "rc=$?"; grep "win bwd" gpurun_out/r2u_prof.log | sort | tail -24 | cut -c1-400
