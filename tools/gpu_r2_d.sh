# stem2 validation + A/B, per-launch DRAM traffic, full ncu capture of every kernel of one step
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "rc bench $?"
tail -3 gpurun_out/bench_r2d.err
YFV2_STEM_FFMA=1 YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2d_ffma.json 2> gpurun_out/bench_r2d_ffma.err; echo "rc bench ffma $?"
timeout 400 python tools/ncu_traffic.py capture gpurun_out/traffic_r2d.csv > gpurun_out/traffic_r2d.log 2>&1; echo "rc traffic $?"
timeout 900 ncu --set full --clock-control none -k 'regex:^(stem|s1c_|s2c_|pw3_|tail_|tc_|decode_nms)' -s 15 -c 15 -o gpurun_out/r2d_full python tools/prof_fwd.py 2 > gpurun_out/ncu_full_r2d.log 2>&1; echo "rc ncu $?"
ls -la gpurun_out/
