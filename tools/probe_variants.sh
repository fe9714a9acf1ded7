#!/bin/bash
# tuning sweep of the streaming probe kernel (prints roofline.frac + kernel_ms per variant / CTAs per SM)
[ $# -gt 0 ] || set -- 0:16 4:16 4:32 4:48 4:64 5:32
for vc in "$@"; do
  v=${vc%%:*}; c=${vc##*:}
  QMS_PROBE_VARIANT=$v QMS_PROBE_CTAS=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-build-metric 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('variant $v ctas $c: probe %.4f ms frac %.3f step %.3f ms' % (r['kernel_ms'], r['frac'], d['ms_per_step']), r['step_ms_breakdown'])"
done
