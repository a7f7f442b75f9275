set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_parity.py > gpurun_out/r02_dp_parity_2gpu_final.log 2>&1; tail -6 gpurun_out/r02_dp_parity_2gpu_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_2gpu_final.json 2> gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/r02_bench_c2_2gpu_final.json; tail -3 gpurun_out/bench_2gpu.err
