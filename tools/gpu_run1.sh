set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/ab_dump.py gpurun_out/ab_new.npz > gpurun_out/ab_new.log 2>&1; echo "rc new $?"
YFV2_S1_OLD=1 timeout 300 python tools/ab_dump.py gpurun_out/ab_old.npz > gpurun_out/ab_old.log 2>&1; echo "rc old $?"
python tools/ab_dump.py --cmp gpurun_out/ab_old.npz gpurun_out/ab_new.npz 2>&1 | tail -15
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -m gpu 2>&1 | tail -5
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2a_new.json 2> gpurun_out/bench_r2a_new.err; echo "rc bench $?"
YFV2_S1_OLD=1 YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2a_old.json 2> gpurun_out/bench_r2a_old.err; echo "rc bench old $?"
tail -3 gpurun_out/ab_new.log gpurun_out/bench_r2a_new.err
rm -f gpurun_out/ab_new.npz gpurun_out/ab_old.npz
