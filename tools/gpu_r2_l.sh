set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_dataloader_cpu.py -m "gpu or not gpu" -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2l_tests.log
