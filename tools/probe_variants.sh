#!/bin/bash
# tuning sweep of the streaming probe kernel (prints roofline.frac + kernel_ms per variant)
for v in 0 1 2 3; do for c in 8 16; do
  QMS_PROBE_VARIANT=$v QMS_PROBE_CTAS=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('variant $v ctas $c: probe %.4f ms frac %.3f step %.3f ms' % (r['kernel_ms'], r['frac'], d['ms_per_step']))"
done; done
