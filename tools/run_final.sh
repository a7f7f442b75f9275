# Final-tree evidence of a round: GPU test suite, smoke, ONE default bench run (~4 GPU-minutes together), then the ncu launch list
# of the headline workload. NOTE: that last step serialises ~6000 kernel launches under ncu and takes ~13 GPU-MINUTES on its
# own - give it its own gpurun call with that much budget left (at the end of round 2 it was cut off by the budget and the
# launch list under profiles/ is the one captured earlier in the round).
set -x
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -rf 2>&1 | tail -25 > gpurun_out/r02_gpu_tests_final.log; tail -4 gpurun_out/r02_gpu_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/r02_bench_default_final.json; tail -2 gpurun_out/bench_final.err
python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/plain_c2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_c2_final.csv python bench.py --steps 2 --warmup 1 --no-per-config --no-cpu --no-fp32-line --no-graph > gpurun_out/ncu_c2_launch.log 2>&1
ls -la gpurun_out/r02_launches_c2_final.csv
