set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2v_tests.log
timeout 400 python tools/umma_check.py > gpurun_out/r2v_umma_check.log 2>&1; echo "check rc=$?"; sed -n 20,34p gpurun_out/r2v_umma_check.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2v_bench.json; tail -3 gpurun_out/r2v_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2v_ncu.log 2>&1; echo "ncu rc=$?"
