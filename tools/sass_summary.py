"""Static evidence from the built library: per kernel, how many Blackwell tensor-core / TMA / tensor-memory instructions its
SASS holds (cuobjdump -sass pvo_b200/libpvo_b200.so), next to registers / shared memory / spills (cuobjdump -res-usage).

    python tools/sass_summary.py > profiles/r01_sass_summary.md

Mnemonics (B200 profiling recipe): tcgen05.mma -> UTC*MMA; tcgen05.ld/st -> LDTM / STTM; TMA -> UTMALDG / UTMASTG / UBLKCP;
legacy mma.sync -> HMMA; cp.async -> LDGSTS."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'pvo_b200', 'libpvo_b200.so')
CLASSES = (('tcgen05.mma', r'UTC\w*MMA'), ('tcgen05.ld/st', r'\b(LDTM|STTM)\b'), ('TMA', r'\b(UTMALDG|UTMASTG|UBLKCP|UTMAPF)\b'),
           ('mbarrier', r'\bSYNCS\b'), ('mma.sync', r'\bHMMA\b'), ('cp.async', r'\bLDGSTS\b'), ('cluster', r'\b(UCGABAR|CGABAR)\w*|\bMAPA\b'),
           ('red/atom', r'\b(RED|ATOM|ATOMG|ATOMS)\b'))


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def short(name):
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    m = re.match(r'(?:void )?([\w:]+(?:<.*>)?)\(', name)
    s = m.group(1) if m else name
    return s if len(s) < 110 else s[:107] + '...'


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.split('\n'):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None or '/*' not in line:
            continue
        counts[cur]['instructions'] += 1 if re.search(r'/\*[0-9a-f]{4}\*/', line) else 0
        for cls, pat in CLASSES:
            if re.search(pat, line):
                counts[cur][cls] += 1
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.split('\n'):
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)), int(re.search(r'LOCAL:(\d+)', line).group(1)) if 'LOCAL:' in line else 0)
    names = demangle(order)
    print('# SASS summary of `pvo_b200/libpvo_b200.so` (sm_100a), `python tools/sass_summary.py`\n')
    print('%d kernels.  Columns: instruction counts by class (static, per kernel), registers per thread, static shared memory, local (spill) bytes.\n' % len(order))
    print('| kernel | instr | ' + ' | '.join(c for c, _ in CLASSES) + ' | regs | smem | local |')
    print('|---|---|' + '---|' * (len(CLASSES) + 3))
    keyf = lambda k: (-(counts[k]['tcgen05.mma'] > 0), -(counts[k]['TMA'] > 0), short(names[k]))
    for k in sorted(order, key=keyf):
        c = counts[k]
        r = usage.get(k, ('', '', ''))
        print('| `%s` | %d | ' % (short(names[k]), c['instructions']) + ' | '.join(str(c[cls]) if c[cls] else '' for cls, _ in CLASSES) + ' | %s | %s | %s |' % r)
    tc = [k for k in order if counts[k]['tcgen05.mma']]
    print('\n%d kernels issue tcgen05.mma (UTC*MMA), %d use TMA, %d use legacy mma.sync (HMMA), %d spill to local memory.'
          % (len(tc), sum(1 for k in order if counts[k]['TMA']), sum(1 for k in order if counts[k]['mma.sync']),
             sum(1 for k in order if usage.get(k, (0, 0, 0))[2] > 0)))


if __name__ == '__main__':
    main()
