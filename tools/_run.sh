mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_integer.py -m gpu -q -x > gpurun_out/r02zd_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02zd_tests.log
timeout 300 python tools/bw_probe.py ccl 2>&1 | tee gpurun_out/r02zd_ccl.log
