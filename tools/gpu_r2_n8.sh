set -x
mkdir -p gpurun_out
NG=${NG:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2n${NG}_bench.json 2> gpurun_out/r2n${NG}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2n${NG}_bench.json; tail -3 gpurun_out/r2n${NG}_bench.err
