set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
YFV2_S1_BIG=1 YFV2_S2_BIG=1 python -m pytest tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_A.err | tail -1 > gpurun_out/bench_A.json
YFV2_BENCH_QUICK=1 YFV2_S1_BIG=1 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_s1big.json
YFV2_BENCH_QUICK=1 YFV2_S2_BIG=1 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_s2big.json
python - <<'PY'
import json
for f in ("bench_A","bench_s1big","bench_s2big"):
    d=json.load(open("gpurun_out/%s.json"%f))
    st={s["stage"]:s["us"] for s in d["stages"]}
    print(f, round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), {k:st[k] for k in ("stem","stage2.0","stage3.0","stage3.1","stage4.0","stage4.1")})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(stem_kernel|tc_|decode_nms)' -s 28 -c 28 --csv --log-file gpurun_out/launches_r1b.csv python scratch/prof_fwd.py 2 > gpurun_out/ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:^stem_kernel -s 1 -c 1 --csv --page raw --log-file gpurun_out/stem_full_r1b.csv python scratch/prof_fwd.py 2 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log
