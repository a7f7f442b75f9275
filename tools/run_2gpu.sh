set -x
for ov in ${OVS:-0 1}; do
DV3_DP_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$ov tools/dp_parity.py > gpurun_out/r02_dp_parity_2gpu_overlap$ov.log 2>&1; tail -4 gpurun_out/r02_dp_parity_2gpu_overlap$ov.log
done
# XL with the B = 16 batch sharded over the 2 ranks: gradient all-reduce after the backward pass vs overlapped with it
for ov in ${OVS:-0 1}; do
DV3_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$ov bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_2gpu_overlap$ov.json 2> gpurun_out/bench_2gpu_overlap$ov.err; python - <<PY
import json
for line in open("gpurun_out/r02_bench_2gpu_overlap$ov.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("overlap $ov weak", d["value"], "strong_xl", {k: d["strong_xl"][k] for k in ("value", "ms_per_step", "allreduce")} if "strong_xl" in d else None)
PY
tail -2 gpurun_out/bench_2gpu_overlap$ov.err
done
