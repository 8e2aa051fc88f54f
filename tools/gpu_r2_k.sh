# register-only wgrad, new stem training kernels: validation, train bench, per-kernel profile
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --mode train --steps 5 > gpurun_out/train1_r2k.json 2> gpurun_out/train1_r2k.err; echo "rc train $?"
tail -3 gpurun_out/train1_r2k.err
timeout 300 python tools/prof_train.py 64 > gpurun_out/prof_train_r2k.json 2> gpurun_out/prof_train_r2k.err; echo "rc proftrain $?"
