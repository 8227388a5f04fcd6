# A/B timing of libbclinear build variants (gpurun_variants/lib_<name>.so) on the decode bench, fp32 and bf16 spectra.
run() { label=$1; sp=$2; shift 2; env "$@" timeout 120 python bench.py --spectrum $sp --no-detail --no-cpu-baseline --steps 2000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RES', sys.argv[1], sys.argv[2], 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3))
    elif 'Error' in l or 'error' in l: print('ERR', sys.argv[1], l.rstrip()[:200])
" "$label" "$sp"; }
for sp in float32 bfloat16; do run base $sp A=1; for v in "$@"; do run $v $sp BCL_LIB_PATH=/root/repo/gpurun_variants/lib_$v.so; done; done
