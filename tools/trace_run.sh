export BP_B200_LIB=tools/ubench/libbp_trace.so BP_TC_TRACE=1
timeout 120 python tools/profile_forward.py --reps 1 2> gpurun_out/trace_load.log >/dev/null
BP_TC_SKIP_LOADS=1 timeout 120 python tools/profile_forward.py --reps 1 2> gpurun_out/trace_skip.log > /dev/null
for f in gpurun_out/trace_load.log gpurun_out/trace_skip.log; do echo $f; python - "$f" <<'PY'
import sys,re
cur=None; last={}
for line in open(sys.argv[1]):
    m=re.match(r'tc_trace layer (\d)',line)
    if m: cur=int(m.group(1)); continue
    m=re.match(r'tile\s+(\d+):\s+(.*)',line)
    if m and cur is not None:
        v=[int(x) for x in m.group(2).split()]
        last[cur]=(int(m.group(1)),v)
for k,(n,v) in sorted(last.items()): print('layer',k,'last tile',n,'finished at cycle',v[7])
PY
done
