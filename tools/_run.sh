mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" > gpurun_out/r02zt_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02zt_tests.log
for v in 4095 0; do SB_GEMM_SMALL_MAX_M=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-propagation --no-e2e --no-roofline > gpurun_out/r02zt_b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02zt_b.json')); print('small_max_m $v:', round(d['value'],3), d['ms_per_step'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02zt_dec_launches.csv python tools/decoder_probe.py --once > /dev/null 2>&1; echo ncu rc=$?
