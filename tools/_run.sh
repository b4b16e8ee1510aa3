mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_amg.py tests/test_gpu_kernels.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02x_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02x_bench.json'))
print(d['value'], d['e2e'], d['config']['phase_ms_per_slice'], d['roofline']['frac'], d['roofline']['encoder']['frac'], d.get('thresholds_open',{}).get('value'))
PY
tail -3 gpurun_out/r02x_bench.err
