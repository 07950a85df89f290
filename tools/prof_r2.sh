set -x
ncu --set full --clock-control none --import-source on -k regex:"cqt_tc_kernel|lognorm_split" --launch-skip 2 -c 2 -o gpurun_out/r2_cqt -f python tools/profile_forward.py --reps 2 > gpurun_out/pf.log 2>&1
ls -la gpurun_out
