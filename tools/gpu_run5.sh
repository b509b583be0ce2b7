set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gru_ -s 8 -c 8 -f -o gpurun_out/gru2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1; echo "ncu3 rc=$?"
