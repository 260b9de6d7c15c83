"""Static SASS evidence per kernel of the shipped library -> profiles/<tag>_sass_evidence.txt

    python profiles/tools/sass_evidence.py r2

Counts, per kernel of ogbench_b200/libogbsampler.so (cuobjdump -sass), the mnemonics that prove the claimed mechanisms:
UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (TMA bulk copy, here shared->global),
LDGSTS = cp.async (16-byte, L1 bypass), SYNCS = mbarrier operations.  Needs no GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PATTERNS = ['UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'ATOMG', 'LDG256', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'IMAD', 'MUFU', 'DMUL', 'DADD', 'BAR']


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
    lib = os.path.join(ROOT, 'ogbench_b200', 'libogbsampler.so')
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r'arch = (sm_\w+)', txt)))
    rows = []
    for f in re.split(r'\n\s*Function : ', txt)[1:]:
        name = f.split('\n', 1)[0].strip()
        c = collections.Counter(m.group(1) for m in re.finditer(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', f))
        c['LDG256'] = len(re.findall(r'LDG\.E[.A-Z0-9]*\.256', f))     # 256-bit global loads (sm_100)
        rows.append((name, sum(v for k, v in c.items() if k != 'LDG256'), c))
    names = subprocess.run(['c++filt'], input='\n'.join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    out = [f'# {os.path.relpath(lib, ROOT)}: cubins for {", ".join(arch)}; static instruction counts per kernel (cuobjdump -sass).',
           '# UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (TMA bulk store shared->global), LDGSTS = cp.async 16 B',
           '# (L1 bypass), SYNCS = mbarrier, ATOMG = global atomics (tile tickets), LDG256 = 256-bit global loads.  Regenerate: python profiles/tools/sass_evidence.py ' + tag,
           '%-74s %6s  %s' % ('kernel', 'instrs', '  '.join(PATTERNS))]
    for (name, n, c), dem in sorted(zip(rows, names), key=lambda r: -r[0][1]):
        dem = re.sub(r'\(.*', '', dem).replace('ogb::', '').replace('(anonymous namespace)::', '').replace('void ', '')
        out.append('%-74s %6d  %s' % (dem[:74], n, '  '.join('%*d' % (len(p), c.get(p, 0)) for p in PATTERNS)))
    path = os.path.join(ROOT, 'profiles', f'{tag}_sass_evidence.txt')
    open(path, 'w').write('\n'.join(out) + '\n')
    print(path, len(rows), 'kernels')


if __name__ == '__main__':
    main()
