# Round-end evidence in one GPU call: warp A/B (v1 / v2 / v3) first -- if the pipelined version misbehaves the rest of the run
# uses the second version --, training tests (same for the tensor-core tconv / narrow convs), full GPU tests, bench line,
# ncu capture of the warp kernel, launch lists.
set -x
mkdir -p gpurun_out
timeout 150 python tools/ab_warp.py > gpurun_out/c4_ab_warp.log 2>&1; rc=$?; cat gpurun_out/c4_ab_warp.log | tail -20
if [ $rc -ne 0 ] || [ "$(grep -c 'v1 - v3| = 0.000e+00' gpurun_out/c4_ab_warp.log)" != "4" ]; then export TECO_WARP_V2=2; echo "WARP_V3_BAD rc=$rc -> TECO_WARP_V2=2"; else echo "WARP_V3_OK"; fi
timeout 300 python -m pytest tests/test_gpu_train.py -q --timeout 120 > gpurun_out/c4_train_tests.log 2>&1; rc=$?; tail -15 gpurun_out/c4_train_tests.log
if [ $rc -ne 0 ]; then export TECO_TRAIN_TC_ALL=0; echo "TRAIN_TC_ALL_BAD rc=$rc -> TECO_TRAIN_TC_ALL=0"; else echo "TRAIN_TC_ALL_OK"; fi
timeout 700 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/c4_gputests.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/c4_gputests.log
timeout 800 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"
tail -10 gpurun_out/c4_bench.err; head -c 300 gpurun_out/c4_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_s2d_v --launch-skip 2 --launch-count 1 -o gpurun_out/r02_warp_v3 python tools/profile_warp.py > gpurun_out/ncu_warp.log 2>&1; echo "ncu warp rc=$?"
TECO_TRAIN_PROFILE=1 TECO_TRAIN_NOGRAPH=1 TECO_TRAIN_PRECISION=bf16 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_frvsr_b.csv python tools/bench_train.py frvsr > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
env | grep TECO_ > gpurun_out/c4_env.txt; cat gpurun_out/c4_env.txt
