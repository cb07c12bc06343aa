#!/bin/bash
# A/B of run-time switches (environment variables) on the current build.  Usage: ab_env.sh "<pytest -k>" "<workloads>" "VAR=val ..." "VAR=val ..." ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
K=$1; WL=$2; shift 2
for E in "$@"; do
  echo "== env: $E"
  env $E timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -1
  for W in $WL; do
    env $E python bench.py --workload $W --steps 40 --warmup 8 --resident-frames 48 --cpu-budget 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'], 'ms/frame',d['ms_per_step'], {k:(v['ms'],v['frac_of_hbm_peak']) for k,v in d['kernels'].items()}, 'enqueue', d['host_enqueue_ms_per_step'])"
  done
done
