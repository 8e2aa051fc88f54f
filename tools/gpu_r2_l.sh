# two-stage wgrad reduction: validation, train bench, per-kernel profile
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --mode train --steps 10 > gpurun_out/train1_r2l.json 2> gpurun_out/train1_r2l.err; echo "rc train $?"
tail -3 gpurun_out/train1_r2l.err
YFV2_TRAIN_WGRAD_TILED=1 timeout 300 python bench.py --mode train --steps 10 > gpurun_out/train1_r2l_tiled.json 2> gpurun_out/train1_r2l_tiled.err; echo "rc train tiled $?"
timeout 300 python tools/prof_train.py 64 > gpurun_out/prof_train_r2l.json 2> gpurun_out/prof_train_r2l.err; echo "rc proftrain $?"
