# full GPU check: tests, per-family stage times, bench line
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -15 gpurun_out/gputest.log
timeout 200 python tools/stage_times.py > gpurun_out/stage.json 2>&1
python - <<PY
import json;d=json.load(open("gpurun_out/stage.json"));print(d["us_per_window"],{k[:12]:v["us_per_window"] for k,v in d["families"].items()})
PY
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"py",d["e2e_python"]["value"],d["e2e_python"]["ms_per_step"],d["e2e_python"]["frac_of_e2e"],d["e2e_python"]["materialized"]["ms_per_step"])
PY
