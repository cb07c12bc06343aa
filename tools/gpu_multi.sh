#!/bin/bash
# multi-GPU visit: halo-exchange parity tests + torchrun bench at N = 1, 2[, 4, 8]
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-m1}; NMAX=${2:-2}
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -12
echo "== pytest multigpu"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -12
for N in 1 2 4 8; do
  [ $N -gt $NMAX ] && break
  for HALO in peer nccl; do
    [ $N -eq 1 ] && [ $HALO = nccl ] && continue
    echo "== bench N=$N halo=$HALO (weak, 1080p band per GPU)"
    OUT=gpurun_out/scale_${TAG}_n${N}_$HALO
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 60 --warmup 10 --cpu-budget 0 > $OUT.json 2> $OUT.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 60 --warmup 10 --halo $HALO > $OUT.json 2> $OUT.err
    fi
    tail -c 1700 $OUT.json; tail -4 $OUT.err
  done
done
for N in 2 4 8; do
  [ $N -gt $NMAX ] && break
  echo "== bench 4k strong N=$N"
  OUT=gpurun_out/scale4k_${TAG}_n$N
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --workload bmfr_taa_4k --steps 40 --warmup 8 --resident-frames 48 > $OUT.json 2> $OUT.err
  tail -c 1500 $OUT.json; tail -4 $OUT.err
done
echo "== done"
