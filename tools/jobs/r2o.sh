mkdir -p gpurun_out
timeout 200 python tools/ncu_targets.py > gpurun_out/r2o_plain_targets.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:attn_|gemm_bf16|layernorm|geglu' -o gpurun_out/r2_targets_full -f python tools/ncu_targets.py > gpurun_out/r2o_ncu_targets.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2o_ncu_targets.log
ls -la gpurun_out/*.ncu-rep
