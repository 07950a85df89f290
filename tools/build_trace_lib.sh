#!/bin/bash
# Builds tools/ubench/libbp_trace.so: the same library compiled with -DBP_TC_TRACE (clock64 pipeline stamps of CTA 0 in the
# conv and CQT kernels, the skip-loads experiment).  Use it through BP_B200_LIB=tools/ubench/libbp_trace.so BP_TC_TRACE=1
# (see tools/trace_run.sh); add -DBP_MBAR_DEBUG to get the wait site of a timed-out mbarrier wait printed.
set -e
cd "$(dirname "$0")/../basic_pitch_b200/csrc"
out=/tmp/bp_trace_build; mkdir -p $out
for f in api hcqt cnn decode tc_conv cqt_tc layout ingest writers; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include -I. -DBP_TC_TRACE "$@" -c $f.cu -o $out/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/ubench/libbp_trace.so $out/*.o -lcudart
ls -la ../../tools/ubench/libbp_trace.so
