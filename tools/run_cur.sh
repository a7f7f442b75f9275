set -x
timeout 300 python -m pytest tests/test_gpu_curiosity.py tests/test_reference_golden.py -q -m gpu --timeout 200 2>&1 | tail -4
timeout 200 python tools/curiosity_prof.py 2>&1 | tail -3 | tee gpurun_out/r02_curiosity_prof.txt
timeout 100 python tools/curiosity_prof.py target > gpurun_out/cur_target.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:disagree --csv --log-file gpurun_out/r02_curiosity_kernels.csv python tools/curiosity_prof.py target > gpurun_out/cur_ncu.log 2>&1
tail -2 gpurun_out/cur_target.log; grep -c disagree gpurun_out/r02_curiosity_kernels.csv
