# One GPU call that refreshes the round's evidence: GPU tests, A/Bs, bench line, ncu of the warp kernel, launch list.
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/c3_gputests.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c3_gputests.log
timeout 200 python tools/ab_warp.py > gpurun_out/c3_ab_warp.log 2>&1; echo "ab_warp rc=$?"; tail -12 gpurun_out/c3_ab_warp.log
timeout 300 python tools/ab_fnet.py 1 3 9 > gpurun_out/c3_ab_fnet.log 2>&1; echo "ab_fnet rc=$?"; tail -5 gpurun_out/c3_ab_fnet.log
for c in 444 592; do timeout 200 python bench.py --headline-only --no-cpu-baseline --clips $c --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('clips', d['config']['clips_per_gpu'], 'value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])"; done > gpurun_out/c3_clips.log 2>&1; cat gpurun_out/c3_clips.log
timeout 800 python bench.py --steps 10 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"
tail -10 gpurun_out/c3_bench.err; head -c 400 gpurun_out/c3_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_s2d_v2 --launch-skip 2 --launch-count 1 -o gpurun_out/r02_warp_v2b python tools/profile_warp.py > gpurun_out/ncu_warp.log 2>&1; echo "ncu warp rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_clip296_b.csv python tools/profile_clip.py > gpurun_out/ncu_clip.log 2>&1; echo "ncu list rc=$?"
TECO_TRAIN_PROFILE=1 TECO_TRAIN_NOGRAPH=1 TECO_TRAIN_PRECISION=bf16 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_frvsr.csv python tools/bench_train.py frvsr > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
