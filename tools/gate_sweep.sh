#!/bin/bash
# tuning sweep of the warp-per-row gate (dense library): CTAs per SM x rows per ticket
[ $# -gt 0 ] || set -- 8:32 8:8 8:128 16:32
for cfg in "$@"; do c=${cfg%%:*}; k=${cfg##*:}
QMS_GATE_CTAS=$c QMS_GATE_CHUNK=$k python tools/scale_run.py proteome --peptides 100000 --spectra 5000 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('gate ctas $c chunk $k', d['match']['spectra_per_s_device'], {k:round(v,2) for k,v in d['match']['per_step'].items() if k.startswith('ms_')})"
done
