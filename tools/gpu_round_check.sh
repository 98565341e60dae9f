# the round's closing GPU check: suite, headline bench line, full-size rounds and a soak of the default traversal against the kd-only one
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_suite.log 2>&1; echo "suite rc $?"
tail -3 gpurun_out/r2z_suite.log
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc $?"
tail -c 300 gpurun_out/r2z_bench_n1.json
timeout 400 python tools/find_bvh_mismatch.py conference > gpurun_out/r2z_mismatch_conference.log 2>&1; echo "conf rc $?"
timeout 400 python tools/find_bvh_mismatch.py dragon-sponza > gpurun_out/r2z_mismatch_dragon.log 2>&1; echo "dragon rc $?"
grep -h "differing pixels" gpurun_out/r2z_mismatch_*.log | cut -c1-400
: > gpurun_out/r2z_soak.log
timeout 200 python tools/soak_bvh_vs_kd.py sponza --rounds 150 --seconds 45 >> gpurun_out/r2z_soak.log 2>&1
timeout 200 python tools/soak_bvh_vs_kd.py cornell --rounds 3000 --seconds 30 >> gpurun_out/r2z_soak.log 2>&1
timeout 200 python tools/soak_bvh_vs_kd.py sibenik --rounds 8 --seconds 45 >> gpurun_out/r2z_soak.log 2>&1
timeout 300 python tools/soak_bvh_vs_kd.py conference --spp 64 --rounds 20 --seconds 45 >> gpurun_out/r2z_soak.log 2>&1
timeout 300 python tools/soak_bvh_vs_kd.py dragon-sponza --spp 32 --rounds 20 --seconds 45 >> gpurun_out/r2z_soak.log 2>&1
cut -c1-400 gpurun_out/r2z_soak.log
