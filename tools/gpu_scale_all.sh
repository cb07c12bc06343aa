#!/bin/bash
# one 8-GPU box: weak scaling N = 1, 2, 4, 8, 4K strong scaling N = 2, 4, 8, 8K strong N = 8 (box time is charged 8-fold)
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-s1}
run() {   # N, name, extra bench args
  local N=$1 OUT=gpurun_out/$2_${TAG}_n$1; shift; shift
  if [ $N -eq 1 ]; then
    timeout 240 python bench.py --gpus 1 --cpu-budget 0 "$@" > $OUT.json 2> $OUT.err
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $OUT.json 2> $OUT.err
  fi
  python - <<EOF
import json
try:
    d=json.loads([l for l in open("$OUT.json") if l.startswith("{")][-1])
    print("$OUT", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "host", d.get("host_enqueue_ms_per_step"), "by rank", d.get("ms_per_step_by_rank"))
except Exception as e:
    print("$OUT FAILED", e)
EOF
}
for N in 1 2 4 8; do run $N scale --steps 60 --warmup 10; done
run 1 scale4k --workload bmfr_taa_4k --steps 40 --warmup 8 --resident-frames 48
for N in 2 4 8; do run $N scale4k --workload bmfr_taa_4k --steps 40 --warmup 8 --resident-frames 48; done
run 8 scale8k --workload bmfr_8k --steps 30 --warmup 6 --resident-frames 36
echo "== done"
