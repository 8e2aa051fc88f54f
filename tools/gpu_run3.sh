set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "rc bench $?"
tail -5 gpurun_out/bench_r2b.err
timeout 300 python bench.py --mode train --steps 5 > gpurun_out/bench_r2b_train1.json 2> gpurun_out/bench_r2b_train1.err; echo "rc train $?"
tail -5 gpurun_out/bench_r2b_train1.err; cat gpurun_out/bench_r2b_train1.json
