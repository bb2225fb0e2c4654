"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump.

  python tools/ncu_lines.py <dump.csv> <kernel substring> [top N] [file substring]
Prints, for the lines of <file> (default trace_kernel.cuh) inside the given kernel: share of the kernel's warp
instructions, share of stall samples, average active lanes, and the source text; then a coarse histogram of the kernel
by source ranges given as extra arguments "name:first-last".
"""
import csv
import sys


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main():
    path, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    fsub = sys.argv[4] if len(sys.argv) > 4 else "trace_kernel.cuh"
    ranges = sys.argv[5:]
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'File Path':
            cur = {'file': r[1], 'rows': [], 'fn': ''}
            blocks.append(cur)
        elif r and r[0] == 'Function Name':
            cur['fn'] = r[1]
        elif cur is not None:
            cur['rows'].append(r)
    total_i = total_s = 0
    lines = []
    for b in blocks:
        if kern not in b['fn']:
            continue
        hdr = b['rows'][0]
        ci = hdr.index('Instructions Executed')
        ct = hdr.index('Thread Instructions Executed')
        cs = hdr.index('# Samples')
        for r in b['rows'][1:]:
            if len(r) <= ct or r[0] == '':
                continue
            i, t, s = num(r[ci]), num(r[ct]), num(r[cs])
            total_i += i
            total_s += s
            if fsub in b['file']:
                lines.append((num(r[0]), r[1], i, t, s))
    print(f"kernel {kern}: {total_i} warp instructions, {total_s} samples")
    for ln, src, i, t, s in sorted(lines, key=lambda d: -d[2])[:top]:
        print(f"  {ln:5d} inst {100 * i / total_i:5.2f}% smp {100 * s / max(total_s, 1):5.2f}% lanes {t / max(i, 1):5.1f}  {src.strip()[:110]}")
    for spec in ranges:
        name, _, rg = spec.partition(':')
        a, b2 = (int(v) for v in rg.split('-'))
        sel = [d for d in lines if a <= d[0] <= b2]
        i = sum(d[2] for d in sel); t = sum(d[3] for d in sel); s = sum(d[4] for d in sel)
        print(f"  [{name:24s}] {a:5d}-{b2:5d} inst {100 * i / total_i:5.1f}% smp {100 * s / max(total_s, 1):5.1f}% lanes {t / max(i, 1):5.1f}")


if __name__ == '__main__':
    main()
