# ncu passes of the round (run under gpurun, one GPU): launch list of the bench command, --set full of the two conv kernels
set -x
cd $GRAFT_REPO_ROOT
US3D_BENCH_PRIME=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_spconv_mt -s 2 -c 1 -o gpurun_out/r2f_conv_mt_full -f python scripts/profile_conv.py > gpurun_out/r2f_ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wgrad -s 2 -c 1 -o gpurun_out/r2f_wgrad_full -f python scripts/profile_conv.py > gpurun_out/r2f_ncu_wgrad.log 2>&1
ncu -i gpurun_out/r2f_conv_mt_full.ncu-rep --page raw --csv > gpurun_out/r2f_conv_mt_full_raw.csv
ncu -i gpurun_out/r2f_wgrad_full.ncu-rep --page raw --csv > gpurun_out/r2f_wgrad_full_raw.csv
tail -3 gpurun_out/r2f_ncu_conv.log gpurun_out/r2f_ncu_wgrad.log
