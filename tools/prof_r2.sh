# final ncu captures of the tensor-core kernels (one chunk of windows per launch)
set -x
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|cqt_ts_kernel|cqt_tc_kernel|lognorm_split" --launch-skip 5 -c 5 -o gpurun_out/r2_final_tc -f python tools/profile_forward.py --reps 2 > gpurun_out/pf.log 2>&1
ls -la gpurun_out | tail -5
