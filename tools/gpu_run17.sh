mkdir -p gpurun_out
timeout 90 python bench.py --layers 16 --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_k16.json 2> gpurun_out/bench_k16.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_k16.json; tail -3 gpurun_out/bench_k16.err
