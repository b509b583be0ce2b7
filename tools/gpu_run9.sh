set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^gemm_kernel -s 27 -c 27 -f -o gpurun_out/ffma python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu5.log 2>&1; echo "ncu5 rc=$?"
