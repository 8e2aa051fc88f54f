# last call of round 2: full GPU parity suite + smoke on the final build; A/B of the rolled-loop register sort in the NMS
set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
YFV2_NMS_SORT_ROLLED=1 timeout 300 python -m pytest tests/test_post_gpu.py -x -q 2>&1 | tail -2
for v in base:X=1 rolled:YFV2_NMS_SORT_ROLLED=1; do
  tag=${v%%:*}; kv=${v#*:}
  env $kv YFV2_BENCH_QUICK=1 timeout 200 python bench.py --steps 10 > gpurun_out/bench_z_$tag.json 2> gpurun_out/bench_z_$tag.err; echo "rc $tag $?"
  python tools/bench_show.py gpurun_out/bench_z_$tag.json | grep -E "value|decode"
done
