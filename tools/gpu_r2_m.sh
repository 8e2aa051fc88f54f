# trainer CUDA graphs, banded 640 heads: validation + benches
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --mode train --steps 20 --warmup 10 > gpurun_out/train1_r2m.json 2> gpurun_out/train1_r2m.err; echo "rc train $?"
tail -3 gpurun_out/train1_r2m.err
YFV2_TRAIN_NOGRAPH=1 timeout 300 python bench.py --mode train --steps 20 --warmup 10 > gpurun_out/train1_r2m_nograph.json 2> gpurun_out/train1_r2m_nograph.err; echo "rc train nograph $?"
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --side 640 --steps 5 > gpurun_out/bench640_r2m.json 2> gpurun_out/bench640_r2m.err; echo "rc 640 $?"
YFV2_HEADS_NOBANDS=1 YFV2_BENCH_QUICK=1 timeout 300 python bench.py --side 640 --steps 5 > gpurun_out/bench640_r2m_old.json 2> gpurun_out/bench640_r2m_old.err; echo "rc 640 old $?"
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo "rc bench $?"
