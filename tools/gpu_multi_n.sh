#!/bin/bash
# N-GPU visit (N = 4 or 8; box time is charged N-fold, so only what needs N GPUs): peer-memory parity test at
# world N without host synchronisation, weak-scaling bench at N, 4K strong-scaling bench at N.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-n4}; N=${2:-4}
nvidia-smi -L | head -8
echo "== pytest multigpu world=$N"
timeout 420 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "$N-peer-async" 2>&1 | tail -6
echo "== bench N=$N halo=peer (weak, 1080p band per GPU)"
OUT=gpurun_out/scale_${TAG}_n${N}_peer
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 60 --warmup 10 > $OUT.json 2> $OUT.err
tail -c 1500 $OUT.json; tail -3 $OUT.err
echo "== bench 4k strong N=$N"
OUT=gpurun_out/scale4k_${TAG}_n$N
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --workload bmfr_taa_4k --steps 40 --warmup 8 --resident-frames 48 > $OUT.json 2> $OUT.err
tail -c 1300 $OUT.json; tail -3 $OUT.err
echo "== done"
