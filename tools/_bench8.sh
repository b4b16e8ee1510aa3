mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02zzq_bench_${N}gpu.json 2> gpurun_out/r02zzq_bench_${N}gpu.err; echo "bench$N rc=$?"; cut -c1-300 gpurun_out/r02zzq_bench_${N}gpu.json; tail -3 gpurun_out/r02zzq_bench_${N}gpu.err; nproc
