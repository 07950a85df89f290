set -x
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 3 -c 3 -o gpurun_out/r2_conv_tc_ts -f python tools/profile_forward.py --reps 2 > gpurun_out/pf.log 2>&1
ls -la gpurun_out
