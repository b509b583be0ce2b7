mkdir -p gpurun_out
timeout 100 python tools/c5_probe.py > gpurun_out/c5_probe.json 2> gpurun_out/c5_probe.err; echo "rc=$?"; cat gpurun_out/c5_probe.json | cut -c1-3000; tail -5 gpurun_out/c5_probe.err
