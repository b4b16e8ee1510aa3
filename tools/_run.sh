mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k gemm > gpurun_out/r02zf_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02zf_tests.log
timeout 200 python tools/gemm_probe.py 0 1 4 5 6 7 15 > /dev/null 2>&1
timeout 200 python tools/gemm_probe.py 2>&1 > gpurun_out/r02zf_gemm_probe.log; grep -A1 "N= 2304\|N= 1728\|N= 1152 K=  288\|N= 4608\|N= 8192" gpurun_out/r02zf_gemm_probe.log | cut -c1-260
timeout 200 python tools/encoder_probe.py > gpurun_out/r02zf_enc.log 2>&1; head -12 gpurun_out/r02zf_enc.log
