# 8-GPU box, second pass (trainer on CUDA graphs): training (configs[2]) at 8 and 1 GPUs
set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --mode train --steps 30 --warmup 10 > gpurun_out/train8g.json 2> gpurun_out/train8g.err; echo "rc train8 $?"
tail -c 900 gpurun_out/train8g.json
sleep 3
timeout 300 $TR --nproc-per-node 4 --master-port 29613 bench.py --gpus 4 --mode train --steps 30 --warmup 10 > gpurun_out/train4g.json 2> gpurun_out/train4g.err; echo "rc train4 $?"
sleep 3
timeout 300 python bench.py --mode train --steps 30 --warmup 10 > gpurun_out/train1g.json 2> gpurun_out/train1g.err; echo "rc train1 $?"
tail -c 400 gpurun_out/train1g.json
