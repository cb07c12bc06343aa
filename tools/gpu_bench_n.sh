#!/bin/bash
# bench.py exactly as the driver launches it, at N GPUs (N = 1: plain python; N > 1: torchrun), both arms optional.
# Usage: gpu_bench_n.sh TAG N [extra bench args...]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=$1; N=$2; shift 2
OUT=gpurun_out/bench_${TAG}_n$N
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 "$@" > $OUT.json 2> $OUT.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $OUT.json 2> $OUT.err
fi
echo "rc=$? wall=${SECONDS}s"; grep -v "^frame #" $OUT.err | grep -i -E "error|Traceback|raise" | head -5
python - <<PY
import json
d=json.loads(open("$OUT.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","banded_equals_single")}, d["config"]["workload"])
print("e2e", d.get("e2e"))
print("kernels", {k:(v["ms"],v["frac_of_hbm_peak"]) for k,v in (d.get("kernels") or {}).items()})
print("cpu", d.get("cpu_baseline"))
print("rank ms", d.get("ms_per_step_by_rank"), "spin", d.get("halo_spin_ms_per_step_by_rank"))
for k,v in (d.get("also") or {}).items():
    print("also", k, {kk:v.get(kk) for kk in ("value","ms_per_step","banded_equals_single","error")}, {kk:(vv["ms"],vv["frac_of_hbm_peak"]) for kk,vv in (v.get("kernels") or {}).items()})
PY
