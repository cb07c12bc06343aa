#!/bin/bash
# The round's evidence visit on ONE GPU: smoke, GPU parity tests, the bench exactly as the driver runs it (both arms),
# the ncu launch list of the same command, ncu --set full captures of every kernel, compute-sanitizer on the small parity
# cases.  Everything lands in gpurun_out/ (scratch); tools/summarize_profile.py turns it into profiles/.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02}
echo "== env"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv,noheader; nproc
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== bench (driver arguments)"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
kill $SMI
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n1.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","host_enqueue_ms_per_step")}, d["config"]["workload"], d["clocks"])
print("e2e", d["e2e"]); print("roofline", {k:d["roofline"].get(k) for k in ("kernel","achieved","frac","traffic")}, d["roofline"].get("issue"))
print("kernels", {k:(v["ms"],v["frac_of_hbm_peak"]) for k,v in d["kernels"].items()}); print("cpu", d["cpu_baseline"])
for k,v in d["also"].items(): print("also", k, v.get("value"), v.get("ms_per_step"), {kk:(vv["ms"],vv["frac_of_hbm_peak"]) for kk,vv in (v.get("kernels") or {}).items()}, v.get("error"))
PY
echo "== reference arm (driver arguments)"
timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; cut -c1-400 gpurun_out/bench_${TAG}_ref.json
echo "== ncu launch list (same command, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --gpus 1 --steps 4 --warmup 3 --no-also --cpu-budget 0 > gpurun_out/ncu_list_$TAG.log 2>&1
grep -c "k_" gpurun_out/launches_$TAG.csv
echo "== ncu --set full"
bash tools/gpu_profile.sh $TAG bmfr_taa_4k k_accumulate k_bmfr_block k_taa | tail -4
bash tools/gpu_profile.sh $TAG bmfr_1080p k_accumulate k_bmfr_block | tail -3
bash tools/gpu_profile.sh $TAG bfr_blend_1080p "k_bfr_block" k_bfr_blend | tail -3
echo "== compute-sanitizer"; bash tools/sanitize.sh 2>&1 | tail -8
echo "== done"
