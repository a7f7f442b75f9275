timeout 300 python tools/graph_split.py c2 15 5 2>&1 | tail -2
DV3_NO_BRANCHES=1 timeout 300 python tools/graph_split.py c2 15 5 2>&1 | tail -2
timeout 300 python tools/graph_split.py c1 15 5 2>&1 | tail -2
DV3_NO_BRANCHES=1 timeout 300 python tools/graph_split.py c1 15 5 2>&1 | tail -2
