mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k gemm > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02t_tests.log
timeout 200 python tools/gemm_probe.py 2>&1 > gpurun_out/r02t_gemm_probe.log; cat gpurun_out/r02t_gemm_probe.log
timeout 200 python tools/encoder_probe.py > gpurun_out/r02t_enc.log 2>&1; head -24 gpurun_out/r02t_enc.log
