set -x
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s1c_kernel -s 2 -c 2 -o gpurun_out/s1c python tools/prof_fwd.py 2 > gpurun_out/ncu_s1c.log 2>&1; echo "rc $?"
tail -3 gpurun_out/ncu_s1c.log
