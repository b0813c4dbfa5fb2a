"""Aggregate the ncu source page (cuda,sass) per CUDA source line and per code region.
usage: ncu_lines.py rep kernel_regex [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = fname = hdr = cur = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': kern = r[1][:60]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    if r[0].isdigit():
        cur = (kern, fname, int(r[0]), r[1].strip()[:100]); agg.setdefault(cur, [0.0, 0.0, 0.0]); continue
    if r[0] != '' or cur is None or len(r) < len(hdr) - 2: continue
    ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); ss = hdr.index('# Samples')
    try:
        a = agg[cur]; a[0] += float(r[ie] or 0); a[1] += float(r[te] or 0); a[2] += float(r[ss] or 0)
    except ValueError:
        pass
for k in sorted(set(c[0] for c in agg)):
    items = [(c, v) for c, v in agg.items() if c[0] == k]
    tot = sum(v[0] for _, v in items); tott = sum(v[1] for _, v in items); tots = sum(v[2] for _, v in items)
    print('=====', k, 'warp-inst %.0f thread-inst %.0f (avg active %.1f) samples %.0f' % (tot, tott, tott / max(tot, 1), tots))
    byfile = {}
    for c, v in items:
        b = byfile.setdefault(c[1], [0, 0]); b[0] += v[0]; b[1] += v[1]
    for f, b in byfile.items():
        print('   file %-32s warp %5.1f%% thread %5.1f%% act %.1f' % (f, 100 * b[0] / tot, 100 * b[1] / tott, b[1] / max(b[0], 1)))
    items.sort(key=lambda cv: -cv[1][0])
    for c, v in items[:top]:
        print('%5.1f%% inst %5.1f%% thr  act=%5.1f | %s:%d %s' % (100 * v[0] / max(tot, 1), 100 * v[1] / max(tott, 1), v[1] / max(v[0], 1), c[1][:12], c[2], c[3]))
