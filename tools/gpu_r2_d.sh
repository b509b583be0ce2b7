set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bench_geometry" > gpurun_out/r2d_geom.log 2>&1; echo "geom rc=$?"; tail -30 gpurun_out/r2d_geom.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
