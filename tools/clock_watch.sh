#!/bin/bash
# usage: clock_watch.sh <outfile> -- samples SM clock/power every 100 ms until killed
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 100 > "$1"
