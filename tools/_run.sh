mkdir -p gpurun_out
export SABER_B200_ALLOW_RANDOM_INIT=1
timeout 900 python -m pytest tests/test_gpu_video.py tests/test_gpu_configs.py tests/test_refstack.py -m gpu -q -x > gpurun_out/r02zi_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02zi_tests.log
timeout 600 python tools/propagation_probe_dist.py 64 1 8 2>&1 | grep "GPU\]" | cut -c1-330
SB_PROP_GRAPH=0 timeout 600 python tools/propagation_probe_dist.py 64 1 8 2>&1 | grep "GPU\]" | cut -c1-330
