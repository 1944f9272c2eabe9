#!/bin/bash
# session-2 call L: per-launch durations of the conv kernels, single-CTA tiles vs CTA pairs
mkdir -p gpurun_out
for v in 0 1; do
  UOC_CONV_2SM=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc -c 80 --csv --log-file gpurun_out/conv_launches_2sm$v.csv python tools/one_frame.py > gpurun_out/ncu_2sm$v.log 2>&1; echo "ncu 2sm=$v exit $?"
done
python - <<'PY'
import csv
for v in (0, 1):
    rows = list(csv.reader(open('gpurun_out/conv_launches_2sm%d.csv' % v)))
    hdr = None; out = []
    for r in rows:
        if 'Kernel Name' in r: hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r)); out.append((d['Kernel Name'][:48], d['Grid Size'], float(d['Metric Value']) / 1000.0))
    half = out[len(out) // 2:]          # second frame (warm)
    print('2sm=%d launches %d total us %.1f' % (v, len(half), sum(o[2] for o in half)))
    for o in half: print('   ', o)
PY
