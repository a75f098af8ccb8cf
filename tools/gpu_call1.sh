set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 120 --durations=15 > gpurun_out/r02b_gputests.log 2>&1; echo "pytest rc=$?"
timeout 800 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -25 gpurun_out/r02b_gputests.log; tail -12 gpurun_out/r02b_bench.err; head -c 1500 gpurun_out/r02b_bench.json
