#!/bin/bash
# one quick visit: GPU parity subset (or "all"), then the 1080p and 4K bench lines reduced to per-kernel times.
# Usage: quick_bench.sh [pytest -k expression | all] 
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
K=${1:-"256x256 or generic or 1080p_chain or combined or rgba16f"}
if [ "$K" = "all" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; else timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -3; fi
python bench.py --workload bmfr_1080p --steps 60 --warmup 10 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1080p ms/frame',d['ms_per_step'], {k:(v['ms'],v['frac_of_hbm_peak']) for k,v in d['kernels'].items()}, 'enqueue', d['host_enqueue_ms_per_step'])"
python bench.py --workload bmfr_taa_4k --steps 30 --warmup 5 --resident-frames 35 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('4k ms/frame',d['ms_per_step'], {k:(v['ms'],v['frac_of_hbm_peak']) for k,v in d['kernels'].items()})"
