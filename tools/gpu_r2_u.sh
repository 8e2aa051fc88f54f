# round 2, session 3, call 6: heads of odd maps on the pixel-pair kernel (default now) vs the one-pixel kernel; pw3 with sibling warps
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -4
YFV2_PW_SIB=1 timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -k "shapes or taps or 352_against" 2>&1 | tail -4
YFV2_PW_SIB=2 timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -k "shapes or taps or 352_against" 2>&1 | tail -4
for v in base:X=1 oddg4:YFV2_HEADS_ODD_G4=1 sib1:YFV2_PW_SIB=1 sib2:YFV2_PW_SIB=2; do
  tag=${v%%:*}; kv=${v#*:}
  env $kv YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_u_$tag.json 2> gpurun_out/bench_u_$tag.err; echo "rc $tag $?"
done
python tools/bench_show.py gpurun_out/bench_u_base.json gpurun_out/bench_u_oddg4.json gpurun_out/bench_u_sib1.json gpurun_out/bench_u_sib2.json
