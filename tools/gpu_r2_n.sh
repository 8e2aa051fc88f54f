# stage3.0 as pw1 + channel-streamed dw->pw: validation + A/B
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo "rc bench $?"
tail -3 gpurun_out/bench_r2n.err
YFV2_S2_48_FUSED=1 YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2n_fused.json 2> gpurun_out/bench_r2n_fused.err; echo "rc bench fused $?"
