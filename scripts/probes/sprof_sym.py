import sys, subprocess, bisect, collections, os
fn = sys.argv[1]
maps = []   # (start, end, offset, path)
samples = []
for line in open(fn):
    if line.startswith('M '):
        p = line[2:].split()
        if len(p) < 6: continue
        a, b = [int(x, 16) for x in p[0].split('-')]
        maps.append((a, b, int(p[2], 16), p[5]))
    elif line.startswith('S'):
        samples.append([int(x, 16) for x in line.split()[1:]])
symtabs = {}
def symtab(path):
    if path in symtabs: return symtabs[path]
    syms = []
    for args in (['nm', '-C', '--defined-only', path], ['nm', '-C', '-D', '--defined-only', path]):
        try:
            out = subprocess.run(args, capture_output=True, text=True).stdout
        except Exception:
            out = ''
        for l in out.splitlines():
            q = l.split(' ', 2)
            if len(q) == 3 and q[1] in 'TtWwVv':
                try: syms.append((int(q[0], 16), q[2]))
                except ValueError: pass
    syms = sorted(set(syms))
    symtabs[path] = ([s[0] for s in syms], [s[1] for s in syms])
    return symtabs[path]
base = {}
for a, b, off, path in maps:
    if path not in base or a - off < base[path]: base[path] = a - off
def resolve(addr):
    for a, b, off, path in maps:
        if a <= addr < b:
            rel = addr - base[path]
            ad, nm_ = symtab(path)
            # non-PIE executables: symbols are absolute
            for r in (rel, addr):
                i = bisect.bisect_right(ad, r) - 1
                if i >= 0 and r - ad[i] < 1 << 20:
                    return nm_[i][:110]
            return os.path.basename(path)
    return '?'
selfc = collections.Counter(); incl = collections.Counter()
cache = {}
for s in samples:
    names = []
    for a in s[2:]:
        if a not in cache: cache[a] = resolve(a - 1 if names else a)
        names.append(cache[a])
    if not names: continue
    selfc[names[0]] += 1
    for n in set(names): incl[n] += 1
N = len(samples)
print('samples', N)
print('--- self')
for n, c in selfc.most_common(25): print('%6.2f%% %s' % (100.0 * c / N, n))
print('--- inclusive')
for n, c in incl.most_common(70): print('%6.2f%% %s' % (100.0 * c / N, n))
