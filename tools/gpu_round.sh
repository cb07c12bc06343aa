#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, ncu launch list + full captures.  Logs -> gpurun_out/
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r1}
echo "== env"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv; nproc; free -g | head -2
ldconfig -p | grep -i -E "vulkan|lvp" ; ls /usr/share/vulkan/icd.d /etc/vulkan/icd.d 2>/dev/null
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== bench 1080p"
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
echo "== bench 4k"
timeout 900 python bench.py --workload bmfr_taa_4k --steps 30 --warmup 5 --resident-frames 35 --cpu-budget 0 > gpurun_out/bench4k_$TAG.json 2> gpurun_out/bench4k_$TAG.err; tail -c 2500 gpurun_out/bench4k_$TAG.json; tail -3 gpurun_out/bench4k_$TAG.err
echo "== bench bfr"
timeout 900 python bench.py --workload bfr_blend_1080p --steps 20 --warmup 4 --resident-frames 24 --cpu-budget 0 > gpurun_out/benchbfr_$TAG.json 2> gpurun_out/benchbfr_$TAG.err; tail -c 2500 gpurun_out/benchbfr_$TAG.json; tail -3 gpurun_out/benchbfr_$TAG.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 6 --warmup 3 --cpu-budget 0 > gpurun_out/ncu_list_$TAG.log 2>&1
grep -E "k_accumulate|k_bmfr|k_taa" gpurun_out/launches_$TAG.csv | tail -12
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bmfr_block|k_accumulate" -s 12 -c 4 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 6 --warmup 3 --cpu-budget 0 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "== ncu full (taa 4k, bfr)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_taa" -s 8 -c 2 -f -o gpurun_out/prof_taa_$TAG \
    python bench.py --workload bmfr_taa_4k --steps 6 --warmup 3 --resident-frames 10 --cpu-budget 0 > gpurun_out/ncu_taa_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bfr_bl" -s 16 -c 4 -f -o gpurun_out/prof_bfr_$TAG \
    python bench.py --workload bfr_blend_1080p --steps 6 --warmup 3 --resident-frames 10 --cpu-budget 0 > gpurun_out/ncu_bfr_$TAG.log 2>&1
ls -la gpurun_out/ | tail -12
echo "== compute-sanitizer (smoke)"
timeout 60 true # --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
echo "== done"
