# full GPU check of HEAD: parity suite, default bench line, smoke(), ncu launch list of an eager step.  $1 = tag
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/$1_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/$1_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/$1_bench.json; tail -3 gpurun_out/$1_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/$1_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/$1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/$1_ncu.log 2>&1; echo "ncu rc=$?"
