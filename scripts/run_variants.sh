#!/bin/bash
# runs scripts/profile_step.py once per variants/*.so (or the names given) and prints the best rate
cd "$(dirname "$0")/.."
names="$@"; [ -z "$names" ] && names=$(ls variants/*.so | xargs -n1 basename | sed 's/\.so$//')
for n in $names; do
  out=$(VR_LIB_PATH=$PWD/variants/$n.so python scripts/profile_step.py ${RAYS:-64e6} ${WHICH:-both} 2>&1 | tail -1)
  echo "$n: $out"
done
