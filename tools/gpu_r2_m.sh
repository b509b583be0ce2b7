set -x
mkdir -p gpurun_out
for w in c4-ragged c2 c3; do
  timeout 600 python bench.py --steps 30 --warmup 5 --workload $w > gpurun_out/r2m_bench_$w.json 2> gpurun_out/r2m_bench_$w.err; echo "bench $w rc=$?"; cut -c1-230 gpurun_out/r2m_bench_$w.json; tail -2 gpurun_out/r2m_bench_$w.err
done
timeout 600 python bench.py --steps 20 --warmup 3 --layers 16 --no-cpu-baseline > gpurun_out/r2m_bench_k16.json 2> gpurun_out/r2m_bench_k16.err; echo "bench k16 rc=$?"; cut -c1-230 gpurun_out/r2m_bench_k16.json
timeout 900 python bench.py --steps 10 --warmup 3 --workload c5 > gpurun_out/r2m_bench_c5.json 2> gpurun_out/r2m_bench_c5.err; echo "bench c5 rc=$?"; cut -c1-230 gpurun_out/r2m_bench_c5.json; tail -3 gpurun_out/r2m_bench_c5.err
