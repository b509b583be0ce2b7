set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gcn_layer_kernel -s 10 -c 3 -o gpurun_out/r2p_gcn_layer -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2p_ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -3 gpurun_out/r2p_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm2_kernel -s 2 -c 2 -o gpurun_out/r2p_gemm2 -f python tools/umma_prof.py > gpurun_out/r2p_ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -3 gpurun_out/r2p_ncu2.log
ls -la gpurun_out/*.ncu-rep
