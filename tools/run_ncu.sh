set -x
python tools/ncu_target.py c3 2 > gpurun_out/ncu_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scan_tc_kernel|wm_losses_kernel|ln_silu_bwd_kernel|ln_silu_fwd_kernel|im2col|col2im' -s 40 -c 40 -o gpurun_out/r02_c3_kernels python tools/ncu_target.py c3 2 > gpurun_out/ncu_c3.log 2>&1
python tools/ncu_target.py c5 1 > gpurun_out/ncu_plain_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scan_tc_kernel' -c 2 -o gpurun_out/r02_c5_scan python tools/ncu_target.py c5 1 > gpurun_out/ncu_c5.log 2>&1
python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/ncu_c2_launch.log 2>&1
python bench.py --workload c4 --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/ncu_c4_launch.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
tail -3 gpurun_out/ncu_c3.log gpurun_out/ncu_c5.log
