timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu --timeout 300 2>&1 | tail -25 > gpurun_out/r2_t3.log; tail -4 gpurun_out/r2_t3.log
for w in c3 c4 L c5; do timeout 300 python bench.py --workload $w --no-per-config --no-cpu --no-fp32-line --steps 5 2>gpurun_out/r2_tc_$w.err > gpurun_out/r2_tc_$w.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_tc_$w.json').read().strip().splitlines()[-1]); print(d['config']['id'], round(d['value'],2), round(d['ms_per_step'],2), 'scan fwd/bwd us', round(d['scan_step_latency_us'],1), round(d['scan_bwd_step_latency_us'],1), {k: v for k, v in d['phases_ms'].items() if v > 0.5})"; tail -2 gpurun_out/r2_tc_$w.err; done
