#!/bin/bash
# Round-end validation on a B200 box (run through gpurun from the repo root):
#   full GPU parity suite, smoke, the bench line (ours + reference arm), the ncu launch list of one step, per-launch DRAM traffic,
#   a full ncu capture of every kernel of one step, and the secondary configs (NMS microbench, batch-1 latency, 640x640, training).
# Outputs land in gpurun_out/; the summaries that are kept go to profiles/ (see profiles/README.md).
set -x
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_$T.err | tail -1 > gpurun_out/bench_$T.json; echo "rc bench $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref_$T.err | tail -1 > gpurun_out/bench_ref_$T.json; echo "rc ref $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(stem|s1c_|s2c_|pw3_|tail_|tc_|decode_nms)' -s 15 -c 15 --csv \
    --log-file gpurun_out/launches_$T.csv python tools/prof_fwd.py 2 > gpurun_out/ncu_launches_$T.log 2>&1; echo "rc launches $?"
timeout 400 python tools/ncu_traffic.py capture gpurun_out/traffic_$T.csv > gpurun_out/traffic_$T.log 2>&1; echo "rc traffic $?"
timeout 900 ncu --set full --clock-control none -k 'regex:^(stem|s1c_|s2c_|pw3_|tail_|tc_|decode_nms)' -s 15 -c 15 -o gpurun_out/full_$T \
    python tools/prof_fwd.py 2 > gpurun_out/ncu_full_$T.log 2>&1; echo "rc ncu full $?"
timeout 300 python tools/nms_phases.py > gpurun_out/nms_phases_$T.json 2> gpurun_out/nms_phases_$T.err; echo "rc phases $?"
timeout 300 python tools/bench_nms.py 10000 256 > gpurun_out/nms_$T.json 2> gpurun_out/nms_$T.err; echo "rc nms $?"
timeout 300 python tools/bench_latency.py > gpurun_out/latency_$T.json 2> gpurun_out/latency_$T.err; echo "rc lat $?"
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --side 640 --steps 5 2> gpurun_out/bench640_$T.err | tail -1 > gpurun_out/bench640_$T.json; echo "rc 640 $?"
timeout 300 python bench.py --mode train --steps 30 --warmup 10 2> gpurun_out/train1_$T.err | tail -1 > gpurun_out/train1_$T.json; echo "rc train $?"
timeout 300 python tools/prof_train.py 64 2> gpurun_out/prof_train_$T.err | tail -1 > gpurun_out/prof_train_$T.json; echo "rc proftrain $?"
