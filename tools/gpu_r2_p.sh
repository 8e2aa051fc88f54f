# round 2, session 3, call 1: NMS phase profile (default / variant 1), heads lane map A/B, parity tests of both
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_post_gpu.py tests/test_forward_gpu.py tests/test_regressions_gpu.py -x -q 2>&1 | tail -3
YFV2_NMS_V=1 timeout 600 python -m pytest tests/test_post_gpu.py -x -q 2>&1 | tail -3
timeout 200 python tools/nms_phases.py > gpurun_out/nms_phases_v0.json 2> gpurun_out/nms_phases_v0.err; cat gpurun_out/nms_phases_v0.json
YFV2_NMS_V=1 timeout 200 python tools/nms_phases.py > gpurun_out/nms_phases_v1.json 2> gpurun_out/nms_phases_v1.err; cat gpurun_out/nms_phases_v1.json
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_p0.json 2> gpurun_out/bench_p0.err; echo "rc $?"
YFV2_BENCH_QUICK=1 YFV2_HEADS_RASTER=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_p_raster.json 2> gpurun_out/bench_p_raster.err; echo "rc $?"
YFV2_BENCH_QUICK=1 YFV2_NMS_V=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_p_nms1.json 2> gpurun_out/bench_p_nms1.err; echo "rc $?"
python tools/bench_show.py gpurun_out/bench_p0.json gpurun_out/bench_p_raster.json gpurun_out/bench_p_nms1.json
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:decode_nms' -s 1 -c 1 -o gpurun_out/nms_full_p python tools/prof_fwd.py 2 > gpurun_out/ncu_nms_p.log 2>&1; echo "rc ncu $?"
