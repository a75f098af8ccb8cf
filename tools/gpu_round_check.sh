# One GPU call that refreshes the round's evidence: GPU tests, bench lines (ours + reference arm), the ncu launch list of the
# metric-config clip batch and `ncu --set full` captures of the dominant tensor kernel and of the warp kernel.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_clip.csv python tools/profile_clip.py > gpurun_out/ncu_clip.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_lin --launch-skip 1 --launch-count 1 -o gpurun_out/r02_conv_lin python tools/profile_trunk.py > gpurun_out/ncu_trunk.log 2>&1; echo "ncu trunk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_s2d --launch-skip 2 --launch-count 1 -o gpurun_out/r02_warp python tools/profile_warp.py > gpurun_out/ncu_warp.log 2>&1; echo "ncu warp rc=$?"
tail -3 gpurun_out/r02_gputests.log; tail -6 gpurun_out/r02_bench.err; head -c 600 gpurun_out/r02_bench.json
