# round 2, session 3, call 12: NMS: dense scan for caps <= 512 (lists behind YFV2_NMS_LISTS), four rows in flight in the unfused scoring pass
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_post_gpu.py -x -q 2>&1 | tail -3
YFV2_NMS_LISTS=1 timeout 600 python -m pytest tests/test_post_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/bench_nms.py 10000 256 > gpurun_out/nms_x.json 2> gpurun_out/nms_x.err; python -c "
import json; d=json.load(open('gpurun_out/nms_x.json'))
for k,v in d['sets'].items(): print(k, v['nms_ms'], v['fused_decode_nms_ms'], v['bit_exact_on_sample'])"
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "rc $?"
python tools/bench_show.py gpurun_out/bench_x.json | tail -3
