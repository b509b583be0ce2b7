set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2w_tests.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2w_bench.json; tail -3 gpurun_out/r2w_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2w_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gcn_layer2_kernel -s 10 -c 3 -o gpurun_out/r2w_gcn_layer2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2w_ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -2 gpurun_out/r2w_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm3_kernel -s 40 -c 12 -o gpurun_out/r2w_gemm3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2w_ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -2 gpurun_out/r2w_ncu2.log
ls -la gpurun_out/*.ncu-rep
