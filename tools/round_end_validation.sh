#!/bin/bash
# Round-end validation on a B200 box (run through gpurun from the repo root):
#   full GPU parity suite, smoke, the bench line, the ncu launch list of one step and the full capture of the dominant kernel.
# Outputs land in gpurun_out/; the summaries that are kept go to profiles/ (see profiles/README.md).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(stem_kernel|tc_|decode_nms)' -s 28 -c 28 --csv \
    --log-file gpurun_out/launches.csv python tools/prof_fwd.py 2 > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:^stem_kernel -s 1 -c 1 --csv --page raw \
    --log-file gpurun_out/stem_full.csv python tools/prof_fwd.py 2 > gpurun_out/ncu_stem.log 2>&1
tail -2 gpurun_out/ncu_stem.log
