"""Per-source-line warp-stall samples of one kernel from a .ncu-rep (ncu --set full --import-source on, -lineinfo build).
    python tools/ncu_lines.py gpurun_out/prof_lbs.ncu-rep k_skin_tc [top_n]
Aggregates the SASS rows of the `cuda,sass` source view under the CUDA line they belong to (first kernel instance only)."""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
fname = "?"
agg = collections.OrderedDict()
cur = None
n_hdr = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        n_hdr += 1
        if n_hdr > 1:
            break
        continue
    if r and r[0] == 'File Name':
        fname = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        i_s = hdr.index('# Samples')
        stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0]:                       # a CUDA source line
        cur = (fname, r[0], r[1].strip()[:90])
        agg.setdefault(cur, [0, collections.Counter()])
    if cur is None:
        continue
    try:
        n = int(r[i_s])
    except ValueError:
        continue
    if r[0]:
        continue                   # the CUDA row repeats the sum of its SASS rows
    agg[cur][0] += n
    for i, h in stalls:
        try:
            agg[cur][1][h] += int(r[i])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
print('total samples', tot)
for (f, ln, src), (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print('%5d %4.1f%%  %s:%s  %s   [%s]' % (n, 100.0 * n / max(tot, 1), f, ln, src, ' '.join('%s=%d' % (k[6:], v) for k, v in st.most_common(3))))
