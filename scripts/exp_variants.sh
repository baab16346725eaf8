#!/bin/bash
# usage: scripts/exp_variants.sh "name[:ENV=V ENV2=V]" ...   (name = variants/<name>.so)
# One line per variant: kernel ms of a C4 trace (RAYS per particle, WHICH particles) and the
# traverse / shade / other split from an extra pass with CUDA events around every launch.
cd "$(dirname "$0")/.."
RAYS=${RAYS:-64e6}; WHICH=${WHICH:-both}
for spec in "$@"; do
  name="${spec%%:*}"; envs=""; [ "$spec" != "$name" ] && envs="${spec#*:}"
  a=$(env $envs VR_LIB_PATH=$PWD/variants/$name.so python scripts/profile_step.py $RAYS $WHICH 2>&1 | tail -1)
  b=$(env $envs VR_TIME_KERNELS=1 VR_LIB_PATH=$PWD/variants/$name.so python scripts/profile_step.py $RAYS $WHICH 2>&1 | grep "phases" | tail -1)
  echo "$spec | $(echo $a | sed 's/.*kernel_ms/kernel_ms/') | $b"
done
