set -x
mkdir -p gpurun_out
date +%s
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; date +%s; cat gpurun_out/bench_n2.json | cut -c1-200; tail -3 gpurun_out/bench_n2.err
