#!/bin/bash
# A/B timing of k_accumulate build variants on the GPU box, then the default build again + smoke + GPU tests + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
CS=vulkanpbrt_b200/csrc
build() {
  nvcc -std=c++17 -O3 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -gencode arch=compute_100a,code=sm_100a $1 -x cu -c $CS/accumulate.cu -o build/obj/accumulate.cu.o 2>&1 | grep -E "error"
  nvcc -shared -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -cudart static -o vulkanpbrt_b200/lib/libvkpbrt_b200.so build/obj/accumulate.cu.o build/obj/taa.cu.o build/obj/bmfr.cu.o build/obj/bfr.cu.o build/obj/halo.cu.o build/obj/api.cpp.o
}
for V in "$@"; do
  build "$V"
  echo "== variant: $V"
  timeout 300 python -m pytest tests/test_parity.py -m gpu -x -q -k "256x256 or 1080p_chain or combined" 2>&1 | tail -1
  python bench.py --steps 60 --warmup 10 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms/frame',d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
  python bench.py --workload bmfr_taa_4k --steps 30 --warmup 5 --resident-frames 35 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   4k ms/frame',d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done
echo "== default build: smoke, GPU tests, bench"
build ""
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 2500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
