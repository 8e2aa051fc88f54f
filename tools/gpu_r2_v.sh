# round 2, session 3, call 8: heads2 epilogue with 8-byte pair stores
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -4
for v in base:X=1; do
  tag=${v%%:*}; kv=${v#*:}
  env $kv YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_w_$tag.json 2> gpurun_out/bench_w_$tag.err; echo "rc $tag $?"
done
python tools/bench_show.py gpurun_out/bench_w_base.json
