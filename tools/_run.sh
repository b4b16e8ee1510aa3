mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "gpupool" > gpurun_out/r02zb_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02zb_tests.log
