#!/bin/bash
# usage: tools/fused_sweep.sh "<shape>" "ENV1=.. ENV2=.." "ENV.." ...   -- one prefill_bench run per env set (fused-launch tuning)
shape=$1; shift
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 120 python tools/prefill_bench.py --shapes $shape --skip-torch --skip-stream 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    print(d['shape'], round(d['gemm_path_us'],1), round(d['gemm_path_hbm_frac'],3), d['gemm_path_status'])
"
done
