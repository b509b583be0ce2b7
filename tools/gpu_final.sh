# What a round-end check runs on a 1-GPU B200 box (under /usr/local/graft/bin/gpurun -- 'bash tools/gpu_final.sh'):
# GPU parity tests, the default bench line, smoke(), and the ncu launch list of an eager (--no-graph) bench run.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/b_ncu.log 2>&1; echo "ncu rc=$?"
