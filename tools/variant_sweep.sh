#!/bin/bash
# tools/variant_sweep.sh OUT "workloads" v1 v2 ... : quick_bench of every library variant (ASGPU_LIB).
out=$1; wl=$2; shift 2
: > gpurun_out/$out
for v in "$@"; do
  echo "== variant $v" | tee -a gpurun_out/$out
  ASGPU_LIB=$PWD/appleseed_b200/libasgpu_$v.so python tools/quick_bench.py --workloads $wl $QB_ARGS 2>/dev/null | grep '^{"workload"' | grep -v wide_nodes | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: v for k, v in d.items() if not k.endswith('_lanes')})
" | tee -a gpurun_out/$out
done
