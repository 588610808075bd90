"""Attribute an ncu capture to source lines / functions of the stepper.

ncu's CSV export of the source page is SASS-only; this joins it (by instruction order) with `nvdisasm --print-line-info`
of the cubin embedded in the shipped library, then sums executed instructions and stall samples per source line and per
enclosing function.

    python profiles/ncu_by_line.py gpurun_out/prof.ncu-rep <kernel substring> [launch ordinal] [--top 25] [--lib path/to/lib.so]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'flybody_b200', 'lib', 'libflybody_b200.so')


def sass_page(rep, ksub, ordinal):
    """-> (kernel name, [(instr executed, thread instr, samples, {stall: n})...]) for the ordinal-th launch whose name contains ksub"""
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'rows': [], 'hdr': None}
            blocks.append(cur)
        elif cur is not None:
            if cur['hdr'] is None:
                cur['hdr'] = row
            else:
                cur['rows'].append(row)
    sel = [b for b in blocks if ksub in b['name']]
    b = sel[ordinal]
    h = b['hdr']
    ie, te, ns = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
    stalls = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
    rows = []
    for r in b['rows']:
        rows.append((int(r[ie] or 0), int(r[te] or 0), int(r[ns] or 0), {n: int(r[i] or 0) for i, n in stalls}, r[1].strip()))
    return b['name'], rows


def cubin_functions(cubin_text):
    """{function: [((file, line), normalised sass text) per instruction]} for every function in the cubin"""
    funcs, cur, loc = {}, None, ('?', 0)
    for line in cubin_text.splitlines():
        m = re.match(r'^\.text\.(\S+):', line)
        if m:
            cur = m.group(1); funcs[cur] = []; loc = ('?', 0); continue
        if cur is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
        if m:
            funcs[cur].append((loc, norm(m.group(1))))
    return funcs


def norm(t):
    return re.sub(r'\s+', ' ', t.replace('`', '').strip().rstrip(';')).split(' ')[0:2].__str__()


def align(rows, funcs):
    """ncu lists the kernel followed by the device functions it calls; align each run of rows with the cubin function
    whose instruction texts match (opcode + first operand of the first instructions)."""
    locs, i = [], 0
    names = list(funcs)
    while i < len(rows):
        best = None
        for n in names:
            f = funcs[n]
            if len(f) <= len(rows) - i and all(norm(rows[i + j][4]) == f[j][1] for j in range(min(len(f), 12))):
                if best is None or len(f) > len(funcs[best]):
                    best = n
        if best is None:
            locs.append(('?', 0)); i += 1; continue
        locs.extend(l for l, _ in funcs[best]); i += len(funcs[best])
    return locs


def function_table():
    """(file, line) -> name of the enclosing top-level function, from a light scan of the sources"""
    table = {}
    src = os.path.join(ROOT, 'flybody_b200', 'csrc')
    for f in os.listdir(src):
        starts = []
        for i, l in enumerate(open(os.path.join(src, f)), 1):
            m = re.match(r'^(?:template\s*<[^>]*>\s*)?(?:FB_DEVN?|FB_WARPFN|static|__global__|inline)[^;(]*?\b(\w+)\s*\(', l)
            if m and not l.startswith(' '):
                starts.append((i, m.group(1)))
        table[f] = starts
    return table


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    lib = sys.argv[sys.argv.index('--lib') + 1] if '--lib' in sys.argv else LIB
    ordinal = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith('--') else 0
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 25
    name, rows = sass_page(rep, ksub, ordinal)
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=td, capture_output=True)
        cub = [os.path.join(td, f) for f in os.listdir(td) if f.endswith('.cubin')][0]
        text = subprocess.run(['nvdisasm', '--print-line-info', cub], capture_output=True, text=True).stdout
    locs = align(rows, cubin_functions(text))
    unmatched = sum(1 for l in locs if l[0] == '?')
    if unmatched:
        print(f'warning: {unmatched} of {len(rows)} instructions not matched to the cubin', file=sys.stderr)
    ftab = function_table()

    def func_of(loc):
        f, ln = loc
        best = '?'
        for s, n in ftab.get(f, []):
            if s <= ln:
                best = n
        return f'{f}:{best}'
    by_line, by_func = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot = [0, 0, 0]
    for (ie, te, ns, st, _), loc in zip(rows, locs):
        for acc in (by_line[loc], by_func[func_of(loc)], tot):
            acc[0] += ie; acc[1] += te; acc[2] += ns
        for k, v in st.items():
            by_func[func_of(loc)][3][k] += v
    print(f'kernel: {name[:110]}')
    print(f'total: {tot[0]} warp instr, {tot[1] / max(tot[0], 1):.2f} threads/instr, {tot[2]} samples')
    print('\n-- by function (warp instr, share, threads/instr, sample share, top stalls)')
    for k, v in sorted(by_func.items(), key=lambda kv: -kv[1][2])[:top]:
        st = sorted(v[3].items(), key=lambda kv: -kv[1])[:3]
        print(f'{k:42s} {v[0]:>11d} {100 * v[0] / tot[0]:5.1f}%  {v[1] / max(v[0], 1):5.1f}  {100 * v[2] / max(tot[2], 1):5.1f}%  '
              + ' '.join(f'{a[6:]}={100 * b / max(v[2], 1):.0f}%' for a, b in st))
    print('\n-- by line')
    for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f'{k[0]}:{k[1]:<6d} {v[0]:>11d} {100 * v[0] / tot[0]:5.1f}%  {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / max(tot[2], 1):5.1f}%')


if __name__ == '__main__':
    main()
