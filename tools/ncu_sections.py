"""Aggregates an `ncu --page source --print-source cuda,sass --csv` dump of the marching kernel by kernel section.

  python tools/ncu_sections.py <dump.csv> <trace_kernel.cuh as profiled>
"""
import csv
import sys


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'File Path':
            cur = {'file': r[1], 'rows': []}
            blocks.append(cur)
        elif r and r[0] == 'Function Name':
            cur['fn'] = r[1]
        elif cur is not None:
            cur['rows'].append(r)
    b = [x for x in blocks if 'trace_kernel' in x.get('fn', '') and x['file'].endswith('trace_kernel.cuh')][0]
    data = []
    for r in b['rows'][1:]:
        if len(r) < 10 or r[0] == '':
            continue
        data.append((num(r[0]), r[1], num(r[4]), num(r[7]), num(r[8])))
    ti = sum(d[3] for d in data)
    ts = sum(d[2] for d in data)
    print('total warp instructions', ti, 'source lines', len(data))
    src = open(sys.argv[2]).read().split('\n')

    def find(pat, start=0):
        for i in range(start, len(src)):
            if pat in src[i]:
                return i + 1
        return 10 ** 6

    marks = [('helpers', 1), ('kernel start', find('trace_kernel(const __grid_constant__')),
             ('REFILL', find('REFILL: idle lanes')), ('FINALIZE', find('FINALIZE: hand')),
             ('record_surface', find('auto record_surface')), ('march_step', find('auto march_step')),
             ('march loop', find('if (P.event_threshold >= 32)')), ('HEAVY (0)', find('(0) leaving the level')),
             ('(1) decide', find('bool do_shade = false;')), ('(2) transmittance', find('bool emit = false;')),
             ('emit hits', find('// emit the hits of this pass')), ('(3)(4b)', find('// (3) Volumetric')),
             ('(5) enter block', find('// (5) recursive_raycast')), ('(6) post', find("// (6) what the event")),
             ('epilogue', find('if (P.debug_warp_times) {', find('// (6) what')))]
    for (n, a), (_, b2) in zip(marks, marks[1:] + [('eof', 10 ** 7)]):
        sel = [d for d in data if a <= d[0] < b2]
        i = sum(d[3] for d in sel)
        t = sum(d[4] for d in sel)
        sm = sum(d[2] for d in sel)
        print(f"{n:18s} {a:5d}-{b2:7d} inst {100 * i / ti:5.1f}% samples {100 * sm / max(ts, 1):5.1f}% lanes {t / max(i, 1):5.1f}")
    sel = [d for d in data if d[0] < marks[1][1]]
    print('helpers, top lines:')
    for d in sorted(sel, key=lambda d: -d[3])[:24]:
        print(f"  {d[0]:4d} {100 * d[3] / ti:4.1f}% lanes {d[4] / max(d[3], 1):4.1f} {d[1][:100]}")


if __name__ == '__main__':
    main()
