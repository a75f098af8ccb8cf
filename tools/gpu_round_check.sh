# One GPU call that refreshes the round's evidence (gpurun -- 'bash tools/gpu_round_check.sh'): GPU tests, warp kernel A/B,
# bench lines (ours + reference arm), ncu launch lists of a metric-config step and of a training step, `ncu --set full`
# captures of the trunk kernel and of the warp kernel.  Outputs land in gpurun_out/; copy what is to be judged into profiles/.
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"
timeout 150 python tools/ab_warp.py > gpurun_out/ab_warp.log 2>&1; echo "ab_warp rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_clip.csv python tools/profile_clip.py > gpurun_out/ncu_clip.log 2>&1; echo "ncu list rc=$?"
TECO_TRAIN_PROFILE=1 TECO_TRAIN_NOGRAPH=1 TECO_TRAIN_PRECISION=bf16 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_frvsr.csv python tools/bench_train.py frvsr > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_lin --launch-skip 1 --launch-count 1 -o gpurun_out/conv_lin python tools/profile_trunk.py > gpurun_out/ncu_trunk.log 2>&1; echo "ncu trunk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_s2d_v2 --launch-skip 2 --launch-count 1 -o gpurun_out/warp_v2 python tools/profile_warp.py > gpurun_out/ncu_warp.log 2>&1; echo "ncu warp rc=$?"
tail -3 gpurun_out/gputests.log; tail -12 gpurun_out/ab_warp.log; tail -9 gpurun_out/bench.err; head -c 600 gpurun_out/bench.json
