#!/bin/bash
# times three layers under each BG_HALO_DEBUG mask
for d in 0 1 2 4 8 9 6 15; do
  echo "== debug mask $d"
  BG_HALO_DEBUG=$d timeout 120 python tools/bench_conv.py 32 256 2>&1 | grep -E "^ (256|128) +(64|32|128) +(64|32|128) " | cut -c1-45
done
