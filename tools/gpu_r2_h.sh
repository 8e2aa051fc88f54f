# NMS per-class kept lists (gated on class diversity), device aug: validation; benches
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "rc bench $?"
tail -3 gpurun_out/bench_r2h.err
timeout 300 python tools/bench_nms.py 10000 256 > gpurun_out/nms_r2h.json 2> gpurun_out/nms_r2h.err; echo "rc nms $?"
