run() { label=$1; shift; env "$@" timeout 120 python bench.py --no-detail --no-cpu-baseline --steps 200 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RES', sys.argv[1], 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3))
    elif 'Error' in l or 'error' in l: print('ERR', sys.argv[1], l.rstrip()[:200])
" "$label"; }
run base A=1
run dbg4 BCL_WS_DEBUG=4
run dbg6 BCL_WS_DEBUG=6
run dbg7 BCL_WS_DEBUG=7
for v in nm8 nm16 st16 st8; do run $v BCL_LIB_PATH=/root/repo/gpurun_variants/lib_$v.so; run ${v}_dbg3 BCL_LIB_PATH=/root/repo/gpurun_variants/lib_$v.so BCL_WS_DEBUG=3; done
