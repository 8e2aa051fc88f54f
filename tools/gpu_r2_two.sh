# 2-GPU sanity run of the bench line (per-rank NUMA binding, barrier + max-over-ranks timing)
set -x
cd $GRAFT_REPO_ROOT
YFV2_BENCH_QUICK=1 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_s3.json 2> gpurun_out/bench_2gpu_s3.err; echo "rc $?"
tail -c 1500 gpurun_out/bench_2gpu_s3.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['config'].get('host'))"
tail -3 gpurun_out/bench_2gpu_s3.err
