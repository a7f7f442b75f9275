set -x
timeout 1800 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_final.log; tail -3 gpurun_out/r02_gpu_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/r02_bench_default_final.json; tail -2 gpurun_out/bench_final.err
python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_c2_final.csv python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/ncu_c2_launch.log 2>&1
python tools/ncu_target.py c2 2 > gpurun_out/ncu_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'observe_scan_cluster_kernel|observe_scan_bwd_cluster_kernel' -c 2 -o gpurun_out/r02_cluster_scans python tools/ncu_target.py c2 2 > gpurun_out/ncu_c2_cluster.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_c2_final.csv; tail -3 gpurun_out/ncu_c2_cluster.log
