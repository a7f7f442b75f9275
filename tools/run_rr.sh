timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu --timeout 300 -k "(c1 or c2) and bf16" -x 2>&1 | tail -25 > gpurun_out/r2_rr_tests.log; tail -25 gpurun_out/r2_rr_tests.log | cut -c1-300
timeout 300 python tools/graph_split.py c2 15 5 2>&1 | tail -2
timeout 300 python tools/graph_split.py c1 15 5 2>&1 | tail -2
for w in c1 c2; do timeout 300 python bench.py --workload $w --no-per-config --no-cpu --no-fp32-line --steps 20 2>gpurun_out/r2_xs_$w.err > gpurun_out/r2_xs_$w.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_xs_$w.json').read().strip().splitlines()[-1]); print(d['config']['id'], round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'scan fwd/bwd us', round(d['scan_step_latency_us'],1), round(d['scan_bwd_step_latency_us'],1))"; tail -2 gpurun_out/r2_xs_$w.err; done
