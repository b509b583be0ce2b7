set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or model or parity or proj or real" > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2t_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --workload c2 > gpurun_out/r2m_bench_c2.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; cut -c1-230 gpurun_out/r2m_bench_c2.json
