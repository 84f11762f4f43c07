set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02u_gputest.log
timeout 900 python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-eager --no-video-row --no-batch-row > gpurun_out/r02u_bench_steps20.json 2>> gpurun_out/r02u_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_ncu_launches.csv python tools/ncu_target.py > gpurun_out/r02u_ncu1.log 2>&1
grep -c conv_tc_kernel gpurun_out/r02b_ncu_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1168 -c 26 -o gpurun_out/r02b_conv_full -f python tools/ncu_target.py > gpurun_out/r02u_ncu2.log 2>&1
ncu -i gpurun_out/r02b_conv_full.ncu-rep --page raw --csv > gpurun_out/r02b_conv_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02b_conv_full_raw.csv gpurun_out/r02b_ncu_full_conv_tc_summary.csv
ls -la gpurun_out/ | tail -12
tail -3 gpurun_out/r02u_gputest.log
