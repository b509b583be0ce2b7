set -x
mkdir -p gpurun_out
NG=${NG:-1}
if [ "$NG" = "1" ]; then
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/r2t_bench_eager.json 2> gpurun_out/r2t_bench.err; echo "bench eager rc=$?"; cut -c1-260 gpurun_out/r2t_bench_eager.json
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2t_bench.json
fi
