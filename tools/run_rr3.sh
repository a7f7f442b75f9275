for w in c1 c2; do for rr in "" 1; do
  DV3_DREAM_RR=$rr timeout 200 python bench.py --workload $w --steps 40 --warmup 8 --no-per-config --no-cpu --no-fp32-line 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$w rr=$rr', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3))"
done; done
