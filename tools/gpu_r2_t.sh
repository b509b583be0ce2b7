set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2t_tests.log
for i in 1 2; do
timeout 120 python bench.py --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/r2t_bench_$i.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2t_bench_$i.json
done
