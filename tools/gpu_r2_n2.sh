set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2n2_bench.json; tail -3 gpurun_out/r2n2_bench.err
timeout 100 python -m pytest tests/test_gpu_dp.py -m gpu -x -q > gpurun_out/r2n2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2n2_tests.log
