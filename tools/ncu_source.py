"""Per-region instruction counts from an .ncu-rep captured with --import-source on (development aid).
Splits the SASS into maximal runs between branch targets and prints the heaviest runs."""
import csv, subprocess, sys, re
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ins = [(int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ismp])) for r in rows[2:] if len(r) > iex and r[ia].startswith('0x')]
base = ins[0][0]
tot = sum(i[2] for i in ins); tots = sum(i[3] for i in ins)
print('total inst', tot, 'samples', tots)
# regions: consecutive instructions with the same executed count (+-0)
reg = []
cur = [ins[0]]
for i in ins[1:]:
    if i[2] == cur[-1][2]: cur.append(i)
    else: reg.append(cur); cur = [i]
reg.append(cur)
big = sorted(reg, key=lambda c: -sum(i[2] for i in c))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for c in sorted(big, key=lambda c: c[0][0]):
    n = sum(i[2] for i in c); s = sum(i[3] for i in c)
    print('%05x-%05x  %4d instrs  exec/instr %9d  total %5.1f%%  samples %5.1f%%   %s' % (c[0][0] - base, c[-1][0] - base, len(c), c[0][2], 100.0 * n / tot, 100.0 * s / tots, c[0][1][:50]))
if len(sys.argv) > 3:
    edges = [int(x, 16) for x in sys.argv[3].split(',')]
    print('buckets:')
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = [i for i in ins if lo <= i[0] - base < hi]
        print('  %05x-%05x  %5d instrs  total %5.1f%%  samples %5.1f%%' % (lo, hi, len(sel), 100.0 * sum(i[2] for i in sel) / tot, 100.0 * sum(i[3] for i in sel) / tots))
