#!/bin/bash
# A/B timing of build variants of ONE kernel source on the GPU box (nvcc is in the image).
# Usage: variants.sh <file.cu> <pytest -k expr> <workload> "<flags1>" "<flags2>" ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
F=$1; K=$2; WL=$3; shift 3
CS=vulkanpbrt_b200/csrc
EXTRA=""; case $F in accumulate.cu|taa.cu|convert.cu) EXTRA="-fmad=false";; esac
for V in "$@"; do
  nvcc -std=c++17 -O3 -lineinfo $EXTRA -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -gencode arch=compute_100a,code=sm_100a $V -x cu -c $CS/$F -o build/obj/$F.o 2>&1 | grep -E "error"
  nvcc -shared -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -cudart static -o vulkanpbrt_b200/lib/libvkpbrt_b200.so build/obj/*.o
  echo "== $F variant: $V"
  timeout 300 python -m pytest tests/test_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -1
  for W in $WL; do
    python bench.py --workload $W --steps 40 --warmup 8 --resident-frames 48 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'], 'ms/frame',d['ms_per_step'], {k:(v['ms'],v['frac_of_hbm_peak']) for k,v in d['kernels'].items()})"
  done
done
