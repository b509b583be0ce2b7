mkdir -p gpurun_out
timeout 120 python tools/umma_rate.py > gpurun_out/r2u_rate.log 2>&1; echo "rc=$?"; cat gpurun_out/r2u_rate.log
