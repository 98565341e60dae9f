set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_suite.log 2>&1; echo "suite rc $?"
tail -3 gpurun_out/r2z_suite.log
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc $?"
tail -c 600 gpurun_out/r2z_bench_n1.json
timeout 400 python tools/find_bvh_mismatch.py conference > gpurun_out/r2z_mismatch_conference.log 2>&1; echo "conf rc $?"
timeout 400 python tools/find_bvh_mismatch.py dragon-sponza > gpurun_out/r2z_mismatch_dragon.log 2>&1; echo "dragon rc $?"
grep -h "differing pixels" gpurun_out/r2z_mismatch_*.log | cut -c1-600
