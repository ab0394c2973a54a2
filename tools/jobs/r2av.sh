mkdir -p gpurun_out
CM3P_LIB_PATH=variants/libproffwd.so timeout 120 python tools/attn_one.py 64 > gpurun_out/r2av_win.log 2>&1; echo "rc=$?"; grep "attn fwd" gpurun_out/r2av_win.log | sort | head -20 | cut -c1-260
CM3P_LIB_PATH=variants/libproffwd.so timeout 120 python tools/attn_one.py -1 > gpurun_out/r2av_glob.log 2>&1; echo "rc=$?"; grep "attn fwd" gpurun_out/r2av_glob.log | sort | head -6 | cut -c1-260
