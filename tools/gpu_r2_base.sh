# round-2 baseline: full GPU suite, bench line, per-launch DRAM traffic, ncu launch list, NMS / latency / 640 / train benches
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "rc bench $?"
tail -3 gpurun_out/bench_r2c.err
timeout 400 python tools/ncu_traffic.py capture gpurun_out/traffic_r2c.csv > gpurun_out/traffic_r2c.log 2>&1; echo "rc traffic $?"
timeout 300 python tools/bench_nms.py 10000 256 > gpurun_out/nms_r2c.json 2> gpurun_out/nms_r2c.err; echo "rc nms $?"
timeout 300 python tools/bench_latency.py > gpurun_out/latency_r2c.json 2> gpurun_out/latency_r2c.err; echo "rc lat $?"
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --side 640 --steps 5 > gpurun_out/bench640_r2c.json 2> gpurun_out/bench640_r2c.err; echo "rc 640 $?"
timeout 300 python bench.py --mode train --steps 5 > gpurun_out/train1_r2c.json 2> gpurun_out/train1_r2c.err; echo "rc train $?"
tail -2 gpurun_out/*_r2c.err
