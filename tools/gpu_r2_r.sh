# round 2, session 3, call 5: NMS with the compacted pair matrix and the redundant 32-bit mask resolve; deploy post-process parity
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_post_gpu.py tests/test_ncnn_post_gpu.py -x -q 2>&1 | tail -15
timeout 200 python tools/nms_phases.py > gpurun_out/nms_phases_t.json 2> gpurun_out/nms_phases_t.err; cat gpurun_out/nms_phases_t.json
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo "rc $?"; tail -3 gpurun_out/bench_t.err
python tools/bench_show.py gpurun_out/bench_t.json
timeout 300 python tools/bench_nms.py 2048 256 > gpurun_out/nms_cfg4_t.json 2> gpurun_out/nms_cfg4_t.err; cat gpurun_out/nms_cfg4_t.json
