# round 2, session 3, call 7: stage4.0 stride-2 rows as LDS.64 + shuffle (default) vs scalar rows; pw3 sibling warps now default
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -4
for v in base:X=1 scalar:YFV2_DWS2_SCALAR=1; do
  tag=${v%%:*}; kv=${v#*:}
  env $kv YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_v_$tag.json 2> gpurun_out/bench_v_$tag.err; echo "rc $tag $?"
done
python tools/bench_show.py gpurun_out/bench_v_base.json gpurun_out/bench_v_scalar.json
