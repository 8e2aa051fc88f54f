# sibling warps (heads2w, s1c<48> SIB=2): validation + A/B
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "rc bench $?"
tail -3 gpurun_out/bench_r2f.err
YFV2_HEADS_OLD=1 YFV2_S1_NOSIB=1 YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2f_old.json 2> gpurun_out/bench_r2f_old.err; echo "rc bench old $?"
timeout 600 ncu --set full --clock-control none -k 'regex:^(tc_head2w|s1c_|s2c_)' -s 5 -c 6 -o gpurun_out/r2f_sib python tools/prof_fwd.py 2 > gpurun_out/ncu_r2f.log 2>&1; echo "rc ncu $?"
