"""Development helper: per-region / per-instruction warp-stall report from an ncu report (source page).
   python tools/stall_report.py <rep> [min_samples]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:  # noqa: BLE001
        return 0.0


def op(r):
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    return m.group(2) if m else '?'


tot = sum(f(r, '# Samples') for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {k: sum(f(r, k) for r in data) for k in reasons}
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f'{k:28s} {v:9.0f} {100 * v / tot:5.1f}%')
for i in range(0, len(data), 64):
    seg = data[i:i + 64]
    s = sum(f(r, '# Samples') for r in seg)
    ex = sum(f(r, 'Instructions Executed') for r in seg)
    if s / tot < 0.002:
        continue
    ops = {}
    for r in seg:
        ops[op(r).split('.')[0]] = ops.get(op(r).split('.')[0], 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    print(i, f'{100 * s / tot:5.1f}%', f'exec {ex / 1e6:7.2f}M', top)
print('--- hot instructions')
for i, r in enumerate(data):
    s = f(r, '# Samples')
    if s >= thr:
        top = sorted(((f(r, k), k) for k in reasons), reverse=True)[:3]
        print(i, r[ix['Source']][:64].ljust(64), int(s), int(f(r, 'Instructions Executed') / 1e3), 'k', [(k[6:], int(v)) for v, k in top if v > 0])
