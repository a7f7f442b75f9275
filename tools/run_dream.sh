timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu --timeout 200 -k "bf16 and (c1 or c2 or c3)" -x 2>&1 | tail -12
for w in c2 c1 c3; do timeout 200 python bench.py --workload $w --no-per-config --no-cpu --no-fp32-line --steps 10 2>gpurun_out/r2_dr_$w.err > gpurun_out/r2_dr_$w.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_dr_$w.json').read().strip().splitlines()[-1]); print(d['config']['id'], round(d['value'],2), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k: v for k, v in d['phases_ms'].items() if k.startswith('dream')})"; tail -2 gpurun_out/r2_dr_$w.err; done
timeout 200 python tools/dream_tc_prof.py c2 2>&1 | tail -6
