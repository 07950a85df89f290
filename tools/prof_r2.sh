set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_start.csv python tools/profile_forward.py --reps 2 > gpurun_out/pf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 3 -c 3 -o gpurun_out/r2_conv_tc_start -f python tools/profile_forward.py --reps 2 >> gpurun_out/pf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cqt_tc_kernel|lognorm_split|decimate" --launch-skip 7 -c 7 -o gpurun_out/r2_front_start -f python tools/profile_forward.py --reps 2 >> gpurun_out/pf.log 2>&1
ls -la gpurun_out
