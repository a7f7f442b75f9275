timeout 200 python tools/dream_rr_prof.py c2 2>&1 | tail -3
timeout 200 python tools/dream_rr_prof.py c1 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu --timeout 300 -k "(c1 or c2) and bf16" -x 2>&1 | tail -25 > gpurun_out/r2_rr_tests.log; tail -4 gpurun_out/r2_rr_tests.log | cut -c1-300
timeout 300 python tools/graph_split.py c2 15 5 2>&1 | tail -2
timeout 300 python tools/graph_split.py c1 15 5 2>&1 | tail -2
