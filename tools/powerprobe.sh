#!/bin/bash
# usage: powerprobe.sh  -> runs each mix for 3 s with nvidia-smi sampling at 100 ms
Q="clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.hw_power_brake_slowdown,temperature.gpu"
for cfg in "0 8" "0 64" "1 64" "2 64" "3 64" "3 8"; do
  set -- $cfg
  nvidia-smi --query-gpu=$Q --format=csv,noheader -lms 100 > gpurun_out/pp_$1_$2.csv &
  SMI=$!
  sleep 0.3
  ./tools/powerprobe $1 $2 3.0
  kill $SMI; wait $SMI 2>/dev/null
  echo "  samples (tail):"; tail -n 12 gpurun_out/pp_$1_$2.csv | sort | uniq -c | sort -rn | head -4
done
