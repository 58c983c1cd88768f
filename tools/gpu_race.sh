#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 16 0" "2 16 0" "1 16 1" "1 32 0"; do
  echo "== qb m ring = $cfg"
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_flat.py $cfg 2>&1 | grep -v "^=========     and" | grep "Race reported\|hazard\|RACECHECK\|^ok\|Error\|Warning" | cut -c1-230 | head -8
done
