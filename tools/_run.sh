mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k gemm > gpurun_out/r02v_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02v_tests.log
timeout 200 python tools/encoder_probe.py > gpurun_out/r02v_enc.log 2>&1; head -24 gpurun_out/r02v_enc.log
timeout 200 python tools/decoder_probe.py > gpurun_out/r02v_dec.log 2>&1; head -12 gpurun_out/r02v_dec.log
