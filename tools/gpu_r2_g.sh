# NMS class skip, channel-streamed stage4.0 (dws2c): validation + A/B; where a training step's time goes
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "rc bench $?"
tail -3 gpurun_out/bench_r2g.err
YFV2_S2_96_OLD=1 YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2g_old.json 2> gpurun_out/bench_r2g_old.err; echo "rc bench old $?"
timeout 300 python tools/bench_nms.py 10000 256 > gpurun_out/nms_r2g.json 2> gpurun_out/nms_r2g.err; echo "rc nms $?"
timeout 300 python tools/prof_train.py 64 > gpurun_out/prof_train_r2g.json 2> gpurun_out/prof_train_r2g.err; echo "rc proftrain $?"
tail -3 gpurun_out/prof_train_r2g.err
timeout 600 ncu --set full --clock-control none -k 'regex:^(tc_dws2c|decode_nms)' -s 2 -c 2 -o gpurun_out/r2g_k python tools/prof_fwd.py 2 > gpurun_out/ncu_r2g.log 2>&1; echo "rc ncu $?"
