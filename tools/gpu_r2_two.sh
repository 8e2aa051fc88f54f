# 2-GPU box: data-parallel gradient check over NCCL, training and inference at 2 GPUs
set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 2 --master-port 29611 tools/ddp_check.py > gpurun_out/ddp_check_2.json 2> gpurun_out/ddp_check_2.err; echo "rc ddp_check $?"
cat gpurun_out/ddp_check_2.json; tail -3 gpurun_out/ddp_check_2.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 300 $TR --nproc-per-node 2 --master-port 29614 bench.py --gpus 2 --mode train --steps 10 > gpurun_out/train2.json 2> gpurun_out/train2.err; echo "rc train2 $?"
grep -E "NVLS|nranks|Connected all|AllReduce|Channel" gpurun_out/train2.err | head -20 > gpurun_out/train2_nccl.txt; tail -c 1500 gpurun_out/train2.json
timeout 300 python bench.py --mode train --steps 10 > gpurun_out/train1.json 2> gpurun_out/train1.err; echo "rc train1 $?"
tail -c 300 gpurun_out/train1.json
rm -f gpurun_out/train2.err
