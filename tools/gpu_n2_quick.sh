#!/bin/bash
# 2-GPU quick visit: halo-exchange parity tests (peer paths) + weak and 4K strong bench lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-q1}
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "2-peer" 2>&1 | tail -2
run() { local OUT=gpurun_out/$1_${TAG}_n2.json; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" 2>/dev/null | tail -1 > $OUT; python -c "
import json
d=json.loads(open('$OUT').read()); print('$OUT', d['ms_per_step'], d['ms_per_step_by_rank'], 'host', d['host_enqueue_ms_per_step'], d['halo_spin_ms_per_step_by_rank'])"; }
run scale --steps 60 --warmup 10
run scale4k --workload bmfr_taa_4k --steps 40 --warmup 8 --resident-frames 48
