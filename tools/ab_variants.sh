# A/B timing of libbclinear build variants (gpurun_variants/lib_<name>.so) on the decode bench.
run() { label=$1; shift; env "$@" timeout 120 python bench.py --no-detail --no-cpu-baseline --steps 300 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RES', sys.argv[1], 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3))
    elif 'Error' in l or 'error' in l: print('ERR', sys.argv[1], l.rstrip()[:200])
" "$label"; }
run base A=1
run base_bf16 A=1 -- 2>/dev/null
for v in "$@"; do run $v BCL_LIB_PATH=/root/repo/gpurun_variants/lib_$v.so; done
