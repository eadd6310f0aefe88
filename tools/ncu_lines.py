"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, re, collections
path = sys.argv[1]
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
hdr = None
agg = collections.OrderedDict()
cur = None
stall_cols = None
for r in rows:
    if 'Instructions Executed' in r:
        hdr = r
        ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        iwf = hdr.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in hdr else None
        iwfi = hdr.index('L1 Wavefronts Shared Ideal') if 'L1 Wavefronts Shared Ideal' in hdr else None
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    first = r[0]
    # cuda source lines have a line number in column 0; sass lines have an address
    if re.fullmatch(r'\d+', first):
        cur = (int(first), r[isrc].strip())
        agg.setdefault(cur, [0, 0, collections.Counter(), 0, 0])
        continue
    if cur is None:
        continue
    try:
        n = int(r[ie]); s = int(r[isamp])
    except ValueError:
        continue
    a = agg[cur]; a[0] += n; a[1] += s
    try:
        if iwf is not None: a[3] += int(r[iwf]); a[4] += int(r[iwfi])
    except ValueError: pass
    for i, h in stall_cols:
        try: a[2][h] += int(r[i])
        except ValueError: pass
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-instructions %d, samples %d" % (tot, ts))
twf = sum(a[3] for a in agg.values())
print('shared wavefronts %d' % twf)
for (ln, text), (n, s, st, wf, wfi) in agg.items():
    if n > tot * thresh / 100 or s > ts * thresh / 100:
        top = ", ".join("%s %d" % (k.replace('stall_', ''), v) for k, v in st.most_common(3))
        print("%5d inst %5.1f%% samp %5.1f%% wf %5.1f%% (x%.1f) [%s] | %s" % (ln, 100 * n / tot, 100 * s / max(ts, 1), 100 * wf / max(twf, 1), wf / max(wfi, 1), top, text[:90]))

# optional phase summary: extra args "name:lo-hi" ...
phases = [a for a in sys.argv[3:] if ':' in a]
if phases:
    print("\nphase summary (share of warp-instructions / of stall samples)")
    for ph in phases:
        name, rng = ph.split(':'); lo, hi = map(int, rng.split('-'))
        n = sum(v[0] for (ln, _), v in agg.items() if lo <= ln <= hi)
        s = sum(v[1] for (ln, _), v in agg.items() if lo <= ln <= hi)
        st = collections.Counter()
        for (ln, _), v in agg.items():
            if lo <= ln <= hi: st.update(v[2])
        top = ", ".join("%s %.0f%%" % (k.replace('stall_', ''), 100 * c / max(s, 1)) for k, c in st.most_common(4))
        w = sum(v[3] for (ln, _), v in agg.items() if lo <= ln <= hi)
        print("%-12s inst %5.1f%%  samples %5.1f%%  wavefronts %5.1f%%  [%s]" % (name, 100 * n / tot, 100 * s / max(ts, 1), 100 * w / max(twf, 1), top))
