# launch list of a short bench run (per-launch device time under ncu: cold-cache + serialised, compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --clips 216 --steps 1 --warmup 3 --python-steps 0 > gpurun_out/b_under_ncu.log 2>&1
tail -2 gpurun_out/b_under_ncu.log | cut -c1-300
