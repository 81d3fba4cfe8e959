for m in 0 1; do MOCO_DEBUG_MODE=$m timeout 100 python tools/gpu_lab.py v1e16_c5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('mode $m', d['case'], round(d['stats_kernel_us'],1))
"; done
