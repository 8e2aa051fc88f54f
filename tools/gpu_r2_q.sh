# round 2, session 3, call 2: NMS with the branch-free IoU front + static 80x3 candidate generation; deploy post-process parity
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_post_gpu.py tests/test_ncnn_post_gpu.py -x -q 2>&1 | tail -15
timeout 200 python tools/nms_phases.py > gpurun_out/nms_phases_q.json 2> gpurun_out/nms_phases_q.err; cat gpurun_out/nms_phases_q.json
YFV2_NMS_GENERIC_CELLS=1 timeout 200 python tools/nms_phases.py > gpurun_out/nms_phases_q_generic.json 2> gpurun_out/nms_phases_q_generic.err; cat gpurun_out/nms_phases_q_generic.json
YFV2_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "rc $?"; tail -3 gpurun_out/bench_q.err
python tools/bench_show.py gpurun_out/bench_q.json
