set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2t_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2t_bench.json; tail -3 gpurun_out/r2t_bench.err
timeout 600 python tools/step_profile.py --graph > gpurun_out/r2t_step_profile.log 2>&1; echo "profile rc=$?"; tail -12 gpurun_out/r2t_step_profile.log
cp gpurun_out/step_timeline.txt gpurun_out/r2t_step_timeline.txt
rm -f gpurun_out/step_trace.json
