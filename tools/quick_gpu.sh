# quick GPU sanity: one small forward + parity subset, under short timeouts
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke.log
