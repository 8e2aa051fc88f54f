# stem2 v2: validation, u8 / f32 bench, full ncu capture of both stem2 instantiations, per-launch DRAM traffic
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "rc bench $?"
tail -3 gpurun_out/bench_r2e.err
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 --input f32 > gpurun_out/bench_r2e_f32.json 2> gpurun_out/bench_r2e_f32.err; echo "rc bench f32 $?"
timeout 400 ncu --set full --clock-control none -k 'regex:^stem2' -s 1 -c 1 -o gpurun_out/r2e_stem_u8 python tools/prof_fwd.py 2 > gpurun_out/ncu_r2e_u8.log 2>&1; echo "rc ncu $?"
YFV2_PROF_F32=1 timeout 400 ncu --set full --clock-control none -k 'regex:^stem2' -s 1 -c 1 -o gpurun_out/r2e_stem_f32 python tools/prof_fwd.py 2 > gpurun_out/ncu_r2e_f32.log 2>&1; echo "rc ncu $?"
timeout 400 python tools/ncu_traffic.py capture gpurun_out/traffic_r2e.csv > gpurun_out/traffic_r2e.log 2>&1; echo "rc traffic $?"
