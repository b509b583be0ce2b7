# 2-GPU checks (gpurun --gpus 2 -- 'bash tools/gpu_n2.sh'): NCCL data-parallel gradient equality + the N=2 bench line.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -x -q > gpurun_out/tests_dp.log 2>&1; echo "dp tests rc=$?"; tail -3 gpurun_out/tests_dp.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
