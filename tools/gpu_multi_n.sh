#!/bin/bash
# N-GPU visit (N = 4 or 8; box time is charged N-fold, so only what needs N GPUs): peer-memory parity test at
# world N without host synchronisation, weak-scaling bench, 4K strong scaling, 8K strong scaling, 4K replicas.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-n4}; N=${2:-4}
nvidia-smi -L | head -8; free -g | head -2
run() {   # name, extra bench args
  local OUT=gpurun_out/$1_${TAG}_n$N; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $OUT.json 2> $OUT.err
  tail -c 1400 $OUT.json; grep -v "^frame #" $OUT.err | tail -2
}
echo "== pytest multigpu world=$N"
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "$N-peer-async" 2>&1 | tail -4
echo "== bench N=$N weak (1080p band per GPU)";      run scale --steps 60 --warmup 10
echo "== bench N=$N 4k chain+taa strong";            run scale4k --workload bmfr_taa_4k --steps 40 --warmup 8 --resident-frames 48
echo "== bench N=$N 8k strong";                      run scale8k --workload bmfr_8k --steps 30 --warmup 6 --resident-frames 36
echo "== bench N=$N 4k replicas";                    run repl4k --workload bmfr_taa_4k --replicas --steps 30 --warmup 6 --resident-frames 36
echo "== done"
