#!/bin/bash
# A/B timing of k_bfr_block build variants on the GPU box.  Usage: bfr_variants.sh "<flags1>" "<flags2>" ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
CS=vulkanpbrt_b200/csrc
i=0
for V in "$@"; do
  i=$((i+1))
  nvcc -std=c++17 -O3 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -gencode arch=compute_100a,code=sm_100a $V -x cu -c $CS/bfr.cu -o build/obj/bfr.cu.o 2>&1 | grep -E "error"
  nvcc -shared -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -cudart static -o vulkanpbrt_b200/lib/libvkpbrt_b200.so build/obj/accumulate.cu.o build/obj/taa.cu.o build/obj/bmfr.cu.o build/obj/bfr.cu.o build/obj/halo.cu.o build/obj/api.cpp.o
  echo "== variant $i: $V"
  timeout 300 python -m pytest tests/test_parity.py -m gpu -x -q -k "bfr" 2>&1 | tail -1
  python bench.py --workload bfr_blend_1080p --steps 20 --warmup 4 --resident-frames 24 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms/frame',d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done
echo "== 4k + taa"
python bench.py --workload bmfr_taa_4k --steps 30 --warmup 5 --resident-frames 35 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   4k ms/frame',d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
python bench.py --steps 60 --warmup 10 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms/frame',d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
