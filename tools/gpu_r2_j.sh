# native trainer + new conv1x1 / BN kernels: validation, train bench, per-kernel profile of a step
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
timeout 300 python bench.py --mode train --steps 5 > gpurun_out/train1_r2j.json 2> gpurun_out/train1_r2j.err; echo "rc train $?"
tail -3 gpurun_out/train1_r2j.err
YFV2_TRAIN_PYOPS=1 YFV2_TRAIN_GEMM_OLD=1 timeout 300 python bench.py --mode train --steps 5 > gpurun_out/train1_r2j_old.json 2> gpurun_out/train1_r2j_old.err; echo "rc train old $?"
timeout 300 python tools/prof_train.py 64 > gpurun_out/prof_train_r2j.json 2> gpurun_out/prof_train_r2j.err; echo "rc proftrain $?"
