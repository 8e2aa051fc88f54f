# NMS: fp32 pre-test, dead rows skipped in the in-chunk pass
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; echo "rc bench $?"
timeout 300 python tools/bench_nms.py 10000 256 > gpurun_out/nms_r2o.json 2> gpurun_out/nms_r2o.err; echo "rc nms $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(decode_nms)' -s 1 -c 1 --csv --log-file gpurun_out/nms_time_r2o.csv python tools/prof_fwd.py 2 > /dev/null 2>&1; tail -2 gpurun_out/nms_time_r2o.csv
