set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/c2_gputests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c2_gputests.log
timeout 200 python tools/ab_warp.py > gpurun_out/c2_ab_warp.log 2>&1; echo "ab_warp rc=$?"; cat gpurun_out/c2_ab_warp.log | tail -12
timeout 300 python tools/ab_tail.py 0 12 24 37 74 > gpurun_out/c2_ab_tail.log 2>&1; echo "ab_tail rc=$?"; cat gpurun_out/c2_ab_tail.log | tail -8
TECO_PROF_T=2 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv3x3_tc_kernel --launch-skip 15 --launch-count 2 -o gpurun_out/r02_tail python tools/profile_clip.py > gpurun_out/ncu_tail.log 2>&1; echo "ncu tail rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_s2d_v2 --launch-skip 2 --launch-count 1 -o gpurun_out/r02_warp_v2 python tools/profile_warp.py > gpurun_out/ncu_warp.log 2>&1; echo "ncu warp rc=$?"
ls -la gpurun_out
