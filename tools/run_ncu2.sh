set -x
timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu --timeout 300 -k "bf16 and (c3 or c5)" 2>&1 | tail -3
python tools/ncu_target.py c3 2 > gpurun_out/ncu_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'wm_losses_kernel|im2col_v4' -s 4 -c 6 -o gpurun_out/r02_c3_losses python tools/ncu_target.py c3 2 > gpurun_out/ncu_c3b.log 2>&1
ls -la gpurun_out/r02_c3_losses.ncu-rep
