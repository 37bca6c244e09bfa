"""Aggregates an `ncu --page source --csv` dump: hottest SASS instructions and stall mix (profiling helper)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
ia, isamp, iexe = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and '(' not in h}
ex = []
seen = set()
iaddr = hdr.index('Address') if 'Address' in hdr else None
for k, r in enumerate(data):
    if iaddr is not None:
        if r[iaddr] in seen:   # the CSV lists a kernel's SASS once per view
            continue
        seen.add(r[iaddr])
    try:
        e = int(r[iexe]); s = int(r[isamp])
    except ValueError:
        continue
    if e > 0:
        ex.append((s, e, k, r))
tot = sum(s for s, _, _, _ in ex)
print('executed static instructions', len(ex), 'dynamic warp-instr', sum(e for _, e, _, _ in ex), 'samples', tot)
mix = {h: sum(int(r[i] or 0) for _, _, _, r in ex) for h, i in stall.items()}
print('stall mix:', {h: round(100.0 * v / max(tot, 1), 1) for h, v in sorted(mix.items(), key=lambda kv: -kv[1]) if v})
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for s, e, k, r in sorted(ex, key=lambda t: -t[0])[:top]:
    reasons = sorted(((int(r[i] or 0), h[6:]) for h, i in stall.items()), reverse=True)[:2]
    print('%6d %5.1f%% exec %9d  #%5d  %-60s %s' % (s, 100.0 * s / tot, e, k, r[ia].strip()[:60], reasons))
