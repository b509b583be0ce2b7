# compute-sanitizer over the kernels added or rewritten in the last session of round 2 (third-generation GEMM incl. operand
# pairs, fused head, nodal attention, LMF, vectorised elementwise kernels)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nograph.py -m gpu -x -q > gpurun_out/r2s_new_tests.log 2>&1; echo "new tests rc=$?"; tail -5 gpurun_out/r2s_new_tests.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests -m gpu -x -q -k "gemm_nt or gemm_nn or gemm_tn or head_and or nodal or lmf" > gpurun_out/r2s_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2s_racecheck.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2s_bench.json
