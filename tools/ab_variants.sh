#!/bin/bash
# A/B timing of kernel variants built by tools/dense_variants.sh: "name:ENV=VALUE,..." per variant ("base" = the library itself)
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  ( if [ "$name" != base ]; then export QMS_B200_LIB=$PWD/pyqms_b200/variants/libqms_b200_$name.so; fi
    for kv in ${envs//,/ }; do export "$kv"; done
    timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('variant [$spec] step ms', round(l['ms_per_step'],2), 'kernel ms', round(l['roofline']['step_ms_breakdown']['ms_gate'],2), 'e2e ms', round(l['e2e']['ms_per_step'],1))
" )
done
