mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "fp32_validation" > gpurun_out/r02zl_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02zl_tests.log
for cfg in "2 4" "4 4" "4 2" "8 2" "1 4"; do set -- $cfg; SB_AMG_BATCH_MULT=$1 SB_GRAPH_LANES=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-propagation --no-e2e > gpurun_out/r02zk_b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02zk_b.json')); print('mult $1 lanes $2:', round(d['value'],3), d['config']['phase_ms_per_slice'])"; done
