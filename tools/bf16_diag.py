"""Diagnostic: full per-tensor report of one forced-draw bf16 parity run (tests/_parity.py) for a bench workload."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from _parity import ParityRun  # noqa: E402
from test_gpu_configs import WORKLOADS, _cfg  # noqa: E402

wid = sys.argv[1]
B, T, H = (int(x) for x in sys.argv[2:5])
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 1
cfg = _cfg(wid, (B, T, H))
run = ParityRun(cfg, steps=1, compute_mode=mode, forced=(mode == 1))
rep = sorted(run.report.items(), key=lambda kv: -kv[1])
print(json.dumps({"wid": wid, "BTH": [B, T, H], "violation": run.forced_violation, "flips": run.forced_flips, "exact": run.exact}))
for k, v in rep[:60]:
    print(f"{v:10.3e}  {k}")
e = run.engine
# run-to-run determinism of the dream in this mode
e.dream_forward(start_from_observe=False)
a = e.tensor("dream.hz").clone()
e.dream_forward(start_from_observe=False)
b = e.tensor("dream.hz").clone()
print("dream rerun max abs diff", float((a - b).abs().max()))
run.close()
