set -x
mkdir -p gpurun_out
timeout 120 python tools/umma_probe_ta.py > gpurun_out/r2r_probe_ta.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2r_probe_ta.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2r_bench.json; tail -3 gpurun_out/r2r_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/r2r_bench_eager.json 2> gpurun_out/r2r_bench_eager.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2r_bench_eager.json
