mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/r02zp_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02zp_tests.log
timeout 300 python tools/win_probe.py 8 2>&1 | tee gpurun_out/r02zp_win_probe.log
