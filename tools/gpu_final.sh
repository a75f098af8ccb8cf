# Round-end evidence in one GPU call: full GPU tests, warp A/B, bench line (ours + reference arm).
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/c5_gputests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c5_gputests.log
timeout 150 python tools/ab_warp.py > gpurun_out/c5_ab_warp.log 2>&1; echo "ab_warp rc=$?"; tail -13 gpurun_out/c5_ab_warp.log
timeout 800 python bench.py --steps 10 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench rc=$?"
tail -10 gpurun_out/c5_bench.err; head -c 300 gpurun_out/c5_bench.json
