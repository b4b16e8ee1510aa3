mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02zzl_bench_2gpu.json 2> gpurun_out/r02zzl_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-400 gpurun_out/r02zzl_bench_2gpu.json; tail -3 gpurun_out/r02zzl_bench_2gpu.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02zzl_bench_ref.json 2> gpurun_out/r02zzl_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02zzl_bench_ref.json
