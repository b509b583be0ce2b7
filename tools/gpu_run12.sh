set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adj_spmm_tc_kernel -c 2 -f -o gpurun_out/spmm_tc_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/b_ncu2.log 2>&1; echo "ncu2 rc=$?"
