"""Hot spots of an .ncu-rep captured with --import-source on (read on the CPU box): the SASS instructions with the
most warp-stall samples, with their dominant stall reasons, and per-opcode totals.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top-n]"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        data.append((r[col['Source']].strip(), int(r[col['# Samples']] or 0), int(r[col['Instructions Executed']] or 0),
                     {h: int(r[col[h]] or 0) for h in stall_cols}))
    total = sum(d[1] for d in data)
    print('total samples', total, ' instructions executed', sum(d[2] for d in data))
    agg = collections.Counter()
    for h in stall_cols:
        agg[h] = sum(d[3][h] for d in data)
    print('stall reasons:', ', '.join('%s %.1f%%' % (h[6:], 100.0 * v / max(total, 1)) for h, v in agg.most_common(8)))
    byop = collections.Counter()
    execop = collections.Counter()
    for src, n, ex, st in data:
        op = re.sub(r'^@!?U?P\d+\s+', '', src).split()[0].split('.')[0] if src else '?'
        byop[op] += n
        execop[op] += ex
    print('samples by opcode:', ', '.join('%s %.1f%%' % (o, 100.0 * v / max(total, 1)) for o, v in byop.most_common(14)))
    print('executed by opcode:', ', '.join('%s %.1f%%' % (o, 100.0 * v / max(sum(execop.values()), 1)) for o, v in execop.most_common(14)))
    idx = sorted(range(len(data)), key=lambda i: -data[i][1])[:top]
    for i in sorted(idx):
        src, n, ex, st = data[i]
        why = ', '.join('%s %d' % (h[6:], v) for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print('%5d %6.2f%%  %-70s %s' % (i, 100.0 * n / max(total, 1), src[:70], why))


if __name__ == '__main__':
    main()
