# round 2, session 3: e2e with the process bound to the GPU-local NUMA node (three runs: variance)
set -x
cd $GRAFT_REPO_ROOT
for i in 1 2 3; do
  YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 20 > gpurun_out/bench_y$i.json 2> gpurun_out/bench_y$i.err; echo "rc $?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_y$i.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['config'].get('host'))"
done
