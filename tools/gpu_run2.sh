set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gru_ -s 8 -c 8 -f -o gpurun_out/gru python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 40 -c 12 -f -o gpurun_out/umma python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu4.log 2>&1; echo "ncu4 rc=$?"
