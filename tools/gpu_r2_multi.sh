# 8-GPU box: training (configs[2]) at 8 / 4 / 1 GPUs with the NCCL log, 640x640 inference (configs[3]) at 8 GPUs
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 300 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --mode train --steps 20 --warmup 10 > gpurun_out/train8.json 2> gpurun_out/train8.err; echo "rc train8 $?"
grep -E "NVLS|nranks=|Connected all|Channel 00/|comm 0x.* rank 0 nranks" gpurun_out/train8.err | head -30 > gpurun_out/train8_nccl.txt; tail -c 900 gpurun_out/train8.json
rm -f gpurun_out/train8.err
sleep 5
timeout 300 $TR --nproc-per-node 4 --master-port 29613 bench.py --gpus 4 --mode train --steps 20 --warmup 10 > gpurun_out/train4.json 2> gpurun_out/train4.err; echo "rc train4 $?"
sleep 5
timeout 300 python bench.py --mode train --steps 20 --warmup 10 > gpurun_out/train1.json 2> gpurun_out/train1.err; echo "rc train1 $?"
tail -c 400 gpurun_out/train1.json
YFV2_BENCH_QUICK=1 timeout 300 $TR --nproc-per-node 8 --master-port 29615 bench.py --gpus 8 --side 640 --steps 5 > gpurun_out/bench640_8.json 2> gpurun_out/bench640_8.err; echo "rc 640x8 $?"
tail -c 600 gpurun_out/bench640_8.json
