"""Lists the loops of a SASS dump (cuobjdump -sass file.o) with their instruction mix: a quick look at what one
iteration of a DP kernel's inner loop costs before spending GPU time.
    python tools/sass_loops.py build/obj/viterbi_profile_q.o [mnemonic-to-rank-by] [--inner]"""
import collections
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    key = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith('--') else None
    txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    ins = []
    for line in txt.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
        elif 'Function :' in line:
            ins.append((-1, line.strip()))
    loops = []
    fn = ''
    start = 0
    for i, (a, t) in enumerate(ins):
        if a == -1:
            fn = t
            start = i + 1
            continue
        if 'BRA' in t:
            m = re.search(r'(0x[0-9a-f]+)\s*$', t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a:
                    j = next((k for k in range(start, i) if ins[k][0] == tgt), None)
                    if j is not None:
                        loops.append((fn, j, i))
    def mix(body):
        c = collections.Counter()
        for _, t in body:
            t = re.sub(r'^@!?U?P\d+\s+', '', t)
            c[t.split()[0].split('.')[0]] += 1
        return c
    rows = []
    if '--inner' in sys.argv:
        loops = [L for L in loops if not any(M is not L and M[0] == L[0] and L[1] <= M[1] and M[2] <= L[2] for M in loops)]
    for fn, j, i in loops:
        c = mix(ins[j:i + 1])
        rows.append((c[key] if key else i - j + 1, fn, j, i, c))
    rows.sort(key=lambda r: -r[0])
    for score, fn, j, i, c in rows[:6]:
        print('%s\n  loop 0x%x..0x%x: %d instructions; %s' % (fn[:120], ins[j][0], ins[i][0], i - j + 1,
                                                             ', '.join('%s %d' % kv for kv in c.most_common())))


if __name__ == '__main__':
    main()
