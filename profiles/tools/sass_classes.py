"""Where a kernel's executed instructions and stall samples go, from an `ncu --set full --import-source on` capture.

    python profiles/tools/sass_classes.py gpurun_out/r2a_c2.ncu-rep [kernel-instance] > profiles/<name>.txt

Three views: (1) instruction mix; (2) instructions grouped by how often they execute per 32-row warp tile -- loops and
per-item / per-output bookkeeping show up as classes; (3) per CUDA source line (needs -lineinfo): executed instructions per
warp tile and share of the stall samples.  Needs no GPU (reads the report)."""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, extra):
    return list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', *extra], capture_output=True, text=True).stdout)))


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = page(rep, [])
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    seg = rows[starts[which]:(starts[which + 1] if which + 1 < len(starts) else len(rows))]
    kernel = seg[0][1]
    hdr = seg[1]
    ci, cs = hdr.index('Instructions Executed'), hdr.index('# Samples')
    ins = []
    for r in seg[2:]:
        try:
            ins.append((r[1].strip(), int(r[ci]), int(r[cs])))
        except (ValueError, IndexError):
            pass
    tot, stot = sum(x[1] for x in ins), sum(x[2] for x in ins)
    per_tile = min((x[1] for x in ins if x[1] > 1000), default=1)
    counts = collections.Counter(x[1] for x in ins if x[1] > 0)
    tile = max((c for c in counts if counts[c] > 50), key=lambda c: counts[c], default=per_tile)   # the straight-line index algebra
    print(f'# {kernel}\n# {rep}: {len(ins)} static instructions, {tot} executed (warp level), {stot} stall samples; one warp tile = 32 rows = {tile} executions')
    print('\n## instruction mix')
    by, bys = collections.Counter(), collections.Counter()
    for s, n, sm in ins:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_\.]+)', s)
        op = m.group(2).split('.')[0] if m else s
        if s.startswith('@!PT'):
            op = '@!PT LDS (filler)'
        by[op] += n
        bys[op] += sm
    for op, n in by.most_common(16):
        print(f'{op:20s} {n / tile:9.1f} /tile {100 * n / tot:5.1f} %   samples {100 * bys[op] / max(stot, 1):5.1f} %')
    print('\n## by execution count per warp tile (a class = the instructions of one loop level)')
    c, s, k = collections.Counter(), collections.Counter(), collections.Counter()
    for _, n, sm in ins:
        c[n] += n
        s[n] += sm
        k[n] += 1
    for n, v in sorted(c.items(), key=lambda kv: -kv[1])[:12]:
        if n:
            print(f'{n / tile:7.2f} x per tile: {k[n]:4d} static instr = {v / tile:7.1f} executed/tile {100 * v / tot:5.1f} %   samples {100 * s[n] / max(stot, 1):5.1f} %')
    print(f'total {tot / tile:.0f} executed instructions per warp tile')
    rows = page(rep, ['--print-source', 'cuda,sass'])
    fname = func = None
    seen = {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif r[0] == 'Function Name':
            func = r[1]
        elif r[0].isdigit() and func == kernel:
            try:
                seen.setdefault((fname, int(r[0])), (r[1].strip(), int(r[6]), int(r[7])))
            except (ValueError, IndexError):
                pass
    if seen:
        st = sum(v[1] for v in seen.values())
        print('\n## by CUDA source line (top 30 by executed instructions)')
        for (f, l), (src, smp, e) in sorted(seen.items(), key=lambda kv: -kv[1][2])[:30]:
            print(f'{f[:20]:20s}:{l:4d} {e / tile:7.1f} /tile  samples {100 * smp / max(st, 1):5.1f} %   {src[:96]}')


if __name__ == '__main__':
    main()
