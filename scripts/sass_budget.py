"""Static instruction budget of the hot loops, from the SASS of the built library (no GPU needed).

    python scripts/sass_budget.py [regex of kernel names] > profiles/<round>_sass_loop_budget.txt

For every kernel whose (demangled) name matches, finds the loops (backward branches), keeps the innermost ones
with a body of at least 60 instructions -- the per-row loops of the streaming kernels, the per-origin loop of the
dense kernel -- and prints their instruction mix.  One pass of a streaming-kernel loop = one lag row of a thread's
group = 4 evaluations; of the dense loop = kSlots (8) evaluations.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mdevaluate_b200', '_cuda', 'libmdevaluate_b200.so')
PIPE = {
    'fp64': ('DADD', 'DFMA', 'DMUL', 'DSETP', 'DMNMX'),
    'xu': ('MUFU', 'F2F', 'I2F', 'F2I', 'FRND', 'I2I'),
    'fma': ('FFMA', 'FFMA2', 'FMUL', 'FMUL2', 'FADD', 'FADD2', 'HFMA2', 'IMAD'),
    'alu': ('LOP3', 'SHF', 'PRMT', 'IADD3', 'VIADD', 'LEA', 'SEL', 'FSEL', 'FMNMX', 'ISETP', 'FSETP', 'MOV', 'UMOV',
            'FSET', 'VOTE', 'VOTEU', 'PLOP3', 'R2UR', 'S2R', 'CS2R', 'LDCU', 'UISETP', 'ULOP3', 'UIADD3', 'USEL'),
    'lsu': ('LDS', 'STS', 'LDG', 'STG', 'LD', 'ST', 'ATOMS', 'ATOMG', 'RED', 'LDC', 'LDSM', 'SYNCS', 'UBLKCP'),
}


def pipe_of(op):
    for name, ops in PIPE.items():
        if op in ops:
            return name
    return 'other'


def kernels():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m and name:
            text = m.group(2).strip()
            ops = re.sub(r'^@!?U?P\w+\s+', '', text).split()
            body.append((int(m.group(1), 16), ops[0].split('.')[0], text))
    if name:
        yield name, body


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def loops(body):
    """(start, end) address ranges of backward branches, innermost first."""
    found = []
    for addr, op, text in body:
        if op == 'BRA':
            m = re.search(r'0x([0-9a-f]+)', text)
            if m and int(m.group(1), 16) < addr:
                found.append((int(m.group(1), 16), addr))
    inner = [r for r in found if not any(o != r and r[0] <= o[0] and o[1] <= r[1] for o in found)]
    return sorted(inner)


def main():
    pattern = re.compile(sys.argv[1] if len(sys.argv) > 1 else r'corr_bulk_kernel<2, (1|3|4|8|11), true>|dense_corr_kernel<true'
                                                                r'|nojump_scan_bulk_kernel<true>')
    all_kernels = list(kernels())
    names = demangle([n for n, _ in all_kernels])
    print('# static instruction mix of the hot loops (cuobjdump -sass of libmdevaluate_b200.so; scripts/sass_budget.py)')
    for mangled, body in all_kernels:
        pretty = re.sub(r'mdb::\(anonymous namespace\)::', '', names[mangled])
        if not pattern.search(pretty):
            continue
        print('\n%s   (%d instructions in total)' % (pretty, len(body)))
        for start, end in loops(body):
            ops = [op for addr, op, _ in body if start <= addr <= end]
            if len(ops) < 60:
                continue
            mix = collections.Counter(ops)
            pipes = collections.Counter()
            for op, n in mix.items():
                pipes[pipe_of(op)] += n
            print('  loop 0x%04x-0x%04x: %d instructions | %s' % (
                start, end, len(ops), '  '.join('%s %d' % (k, pipes[k]) for k in ('fp64', 'xu', 'fma', 'alu', 'lsu', 'other'))))
            print('    ' + ' '.join('%s:%d' % kv for kv in mix.most_common()))


if __name__ == '__main__':
    main()
