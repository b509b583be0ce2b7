set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2w_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2w_bench.json; tail -3 gpurun_out/r2w_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2w_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
timeout 200 python tools/step_profile.py --graph > gpurun_out/r2w_prof.log 2>&1; echo "prof rc=$?"; tail -12 gpurun_out/r2w_prof.log; rm -f gpurun_out/step_trace.json
