run() { for w in c2 c3 c4 c5; do timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-per-config --no-cpu --no-fp32-line 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $w', round(d['value'],2), 'ms', round(d['ms_per_step'],3))"; done; }
python tools/ln_probe.py 2>&1 | tail -22
run new
cp dreamer_v3_b200/libdv3.so /tmp/new.so; cp dreamer_v3_b200/libdv3_old.so dreamer_v3_b200/libdv3.so
run old
cp /tmp/new.so dreamer_v3_b200/libdv3.so
