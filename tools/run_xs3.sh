timeout 200 python tools/bwd_cluster_check.py c2 gpurun_out/g_new.npy 2>&1 | tail -2
DV3_NO_CLUSTER_BWD=1 timeout 200 python tools/bwd_cluster_check.py c2 gpurun_out/g_old.npy 2>&1 | tail -2
timeout 100 python tools/bwd_cluster_check.py --cmp gpurun_out/g_old.npy gpurun_out/g_new.npy 2>&1 | tail -20
timeout 200 python tools/scan_stamps.py c2 1 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_engine.py tests/test_gpu_ops.py -q -m gpu --timeout 300 -k "c1 or c2 or cluster or repr or bf16" 2>&1 | tail -15 > gpurun_out/r2_xs3_tests.log; tail -3 gpurun_out/r2_xs3_tests.log
for w in c1 c2; do timeout 300 python bench.py --workload $w --no-per-config --no-cpu --no-fp32-line --steps 20 2>gpurun_out/r2_xs_$w.err > gpurun_out/r2_xs_$w.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_xs_$w.json').read().strip().splitlines()[-1]); print(d['config']['id'], round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'scan fwd/bwd us', round(d['scan_step_latency_us'],1), round(d['scan_bwd_step_latency_us'],1))"; tail -2 gpurun_out/r2_xs_$w.err; done
rm -f gpurun_out/g_new.npy gpurun_out/g_old.npy
