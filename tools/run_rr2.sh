# row-resident dream kernel (opt-in): per-phase profile, its parity tests, graph-replay time of the actor / critic step with and without it
timeout 200 python tools/dream_rr_prof.py c2 2>&1 | tail -3
timeout 200 python tools/dream_rr_prof.py c1 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu --timeout 300 -k "opt_in" 2>&1 | tail -4
for v in "" 1; do DV3_DREAM_RR=$v timeout 300 python tools/graph_split.py c2 15 5 2>&1 | tail -2; DV3_DREAM_RR=$v timeout 300 python tools/graph_split.py c1 15 5 2>&1 | tail -2; done
