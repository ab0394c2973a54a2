mkdir -p gpurun_out
for lib in cm3p_b200/libcm3p_b200.so variants/libs5.so; do echo "== $lib"; CM3P_LIB_PATH=$lib timeout 300 python tools/bench_kernels.py 2>&1 | grep gemm | cut -c1-140; CM3P_LIB_PATH=$lib timeout 300 python tools/bench_kernels.py gemmbwd 2>&1 | grep kernel | cut -c1-140; done
