# native trainer: validation, train bench native vs op-by-op, where the step's time goes; NMS lists without extra shared memory
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --mode train --steps 5 > gpurun_out/train1_r2i.json 2> gpurun_out/train1_r2i.err; echo "rc train $?"
tail -3 gpurun_out/train1_r2i.err
YFV2_TRAIN_PYOPS=1 timeout 300 python bench.py --mode train --steps 5 > gpurun_out/train1_r2i_pyops.json 2> gpurun_out/train1_r2i_pyops.err; echo "rc train pyops $?"
timeout 300 python tools/prof_train.py 64 > gpurun_out/prof_train_r2i.json 2> gpurun_out/prof_train_r2i.err; echo "rc proftrain $?"
timeout 300 python tools/bench_nms.py 10000 256 > gpurun_out/nms_r2i.json 2> gpurun_out/nms_r2i.err; echo "rc nms $?"
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "rc bench $?"
