mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "fp32_validation" > gpurun_out/r02zn_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r02zn_tests.log
