mkdir -p gpurun_out
export SABER_B200_ALLOW_RANDOM_INIT=1
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
SB_PROP_SHARD=zslab timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/propagation_probe_dist.py 64 1 8 > gpurun_out/r02za_prop_2gpu_zslab.log 2>&1; grep "GPU\]" gpurun_out/r02za_prop_2gpu_zslab.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02za_bench_2gpu.json 2> gpurun_out/r02za_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r02za_bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r02za_bench_2gpu.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d.get('propagation'), indent=1))"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02za_bench_1gpu.json 2> gpurun_out/r02za_bench_1gpu.err; echo "bench1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02za_bench_1gpu.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d.get('propagation'), indent=1))"
