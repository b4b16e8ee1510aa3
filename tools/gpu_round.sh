#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of one bench step, full capture of the GEMM.
# usage: tools/gpu_round.sh <tag> [tests|bench|launches|full ...]
TAG=${1:-run}; shift
WHAT=${@:-tests bench launches}
mkdir -p gpurun_out
for w in $WHAT; do
case $w in
tests) timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log;;
smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log;;
bench) SB_GEMM_SHAPES=gpurun_out/${TAG}_gemm_shapes.txt timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err;;
# NB: ncu replays every kernel of the CUDA-graph lanes serially (~50 ms each): a whole bench step (26 k launches) costs ~25
# GPU-minutes. One lane, no warm-up step and a launch cap keep this under ~6 minutes; the list then covers model build,
# graph capture and the first part of the slice (shares are what matter).
launches) SB_GRAPH_LANES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-roofline --no-propagation > gpurun_out/${TAG}_launches.log 2>&1; echo "launches rc=$?"; wc -l gpurun_out/${TAG}_launches.csv;;
full) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3000 -c 3 -o gpurun_out/${TAG}_gemm python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline --no-propagation > gpurun_out/${TAG}_full.log 2>&1; echo "full rc=$?";;
gprobe) timeout 300 python tools/gemm_probe.py > gpurun_out/${TAG}_gemm_probe.log 2>&1; cat gpurun_out/${TAG}_gemm_probe.log;;
gncu) timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 2 -o gpurun_out/${TAG}_gemm1 python tools/gemm_probe.py 0 > gpurun_out/${TAG}_gncu.log 2>&1; echo "gncu rc=$?"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 2 -o gpurun_out/${TAG}_gemm2 python tools/gemm_probe.py 11 >> gpurun_out/${TAG}_gncu.log 2>&1;;
dprobe) timeout 600 python tools/decoder_probe.py > gpurun_out/${TAG}_decoder_probe.log 2>&1; cat gpurun_out/${TAG}_decoder_probe.log; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_dec_launches.csv python tools/decoder_probe.py --once > gpurun_out/${TAG}_dprobe_ncu.log 2>&1; echo "dprobe ncu rc=$?";;
eprobe) timeout 600 python tools/encoder_probe.py > gpurun_out/${TAG}_encoder_probe.log 2>&1; cat gpurun_out/${TAG}_encoder_probe.log; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_enc_launches.csv python tools/encoder_probe.py --once > gpurun_out/${TAG}_eprobe_ncu.log 2>&1; echo "eprobe ncu rc=$?";;
pprobe) timeout 900 python tools/propagation_probe.py 32 1 8 > gpurun_out/${TAG}_propagation_probe.log 2>&1; cat gpurun_out/${TAG}_propagation_probe.log;;
dfull) timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"i2t_tc|t2i_tc" -s 5 -c 3 -o gpurun_out/${TAG}_dec_fused python tools/decoder_probe.py --once > gpurun_out/${TAG}_dfull.log 2>&1; echo "dfull rc=$?";;
efull) timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"flash_attn" -s 14 -c 2 -o gpurun_out/${TAG}_enc_attn python tools/encoder_probe.py --once > gpurun_out/${TAG}_efull.log 2>&1; echo "efull rc=$?";;
diag) timeout 900 python tools/kernel_diag.py > gpurun_out/${TAG}_kernel_diag.log 2>&1; tail -40 gpurun_out/${TAG}_kernel_diag.log;;
esac
done
