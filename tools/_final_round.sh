tools/gpu_round.sh r02zzi tests bench
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_lines -c 4 -o gpurun_out/r02zzi_fft python tools/next_probe.py fftonce > gpurun_out/r02zzi_fft_ncu.log 2>&1; echo "ncu fft rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02zzi_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02zzi_smoke.log
