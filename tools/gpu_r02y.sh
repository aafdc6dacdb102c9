#!/usr/bin/env bash
# r02_y: where the Croatian map loop's 1.5 ms went (6.3 -> 7.8 ms between r02_m and r02_x): the map loops under the new switches.
set -u
mkdir -p gpurun_out
O=gpurun_out/y_maploop.txt
: > $O
for c in c3 c4; do
  python tools/maploop_probe.py $c >> $O 2>> gpurun_out/y_err.log
  RFSI_B200_RANK_CACHE=0 python tools/maploop_probe.py $c >> $O 2>> gpurun_out/y_err.log
  RFSI_B200_RANK_BUCKETS=4096 python tools/maploop_probe.py $c >> $O 2>> gpurun_out/y_err.log
  RFSI_B200_RANK_BUCKETS=4096 RFSI_B200_RANK_CACHE=0 python tools/maploop_probe.py $c >> $O 2>> gpurun_out/y_err.log
  RFSI_B200_RANK_BUCKETS=65536 python tools/maploop_probe.py $c >> $O 2>> gpurun_out/y_err.log
done
RFSI_B200_TRACE=1 python tools/maploop_probe.py c3 > gpurun_out/y_trace.txt 2>&1
cat $O; tail -n 30 gpurun_out/y_trace.txt
