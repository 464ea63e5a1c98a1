#!/bin/bash
# randomized parity stress of the bracket path: (dtype, distribution, n, interp, qs) x seeds, prints only failures
DEC=0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8,0.9
for cfg in "float64 normal 1048583 linear $DEC" "float32 lognormal 524309 midpoint 0.5,0.99,0.01" "float64 uniform 3000017 linear $DEC" \
           "int64 normal 700001 lower 0.25,0.5,0.75" "uint32 ties 400009 higher 0.1,0.5,0.9" "int32 clustered 300007 nearest 0.001,0.999,0.5" \
           "float32 rounded 900001 linear 0.33,0.66" "float64 sorted 2000003 lower 0,1,0.5,0.123" "uint64 uniform 1000003 higher 0.5,0.9999" \
           "float32 sorted 70001 lower 0.1,0.5,0.9" "float64 reversed 300007 nearest 0.25,0.75" \
           "float64 extremes 300001 lower 0.2,0.8" "float32 clustered 1200007 midpoint 0.5" "int64 reversed 800003 nearest 0.05,0.95"; do
  set -- $cfg
  echo "== $cfg"; timeout 300 python tools/dbg_bracket.py $1 $2 $3 $4 $5 ${SEEDS:-12} 2>&1 | grep -E "^seed|^bad|Error|error" | tail -4
done
