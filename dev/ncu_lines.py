"""Aggregate ncu SASS-level stall samples per CUDA source line (ncu --page source --csv + nvdisasm -g line table)."""
import csv, re, sys, collections
sass_csv, disasm, pat = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
# line table per function from nvdisasm
tables = {}; fn = None; cur = None
for line in open(disasm):
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m: fn = m.group(1); tables[fn] = {}; cur = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*)', line)
    if m and fn: tables[fn][int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
blocks = []; b = None
for r in rows:
    if r and r[0] == 'Kernel Name': b = {'name': r[1], 'hdr': None, 'rows': []}; blocks.append(b); continue
    if b is None: continue
    if b['hdr'] is None: b['hdr'] = r; continue
    b['rows'].append(r)
for b in blocks:
    if pat not in b['name']: continue
    h = b['hdr']; ai = h.index('Address'); si = h.index('# Samples'); ii = h.index('Instructions Executed')
    stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    # find matching function table: by template args
    key = re.sub(r'\(int\)', '', b['name'])
    nums = re.findall(r'<([^>]*)>', key)
    cand = [f for f in tables if 'fused' in f or 'k_qp' in f or 'warp' in f]
    want = 'ILi' + 'ELi'.join(x.strip() for x in nums[0].split(',')) + 'E' if nums else ''
    fnm = [f for f in cand if want in f]
    tab = tables[fnm[0]] if fnm else {}
    base = int(b['rows'][0][ai], 16)
    per = collections.Counter(); inst = collections.Counter(); stalls = collections.defaultdict(collections.Counter)
    tot = 0
    for r in b['rows']:
        off = int(r[ai], 16) - base
        ln = tab.get(off, ('?', off))
        n = int(r[si]); per[ln] += n; tot += n; inst[ln] += int(r[ii])
        for c in stall_cols: stalls[ln][h[c]] += int(r[c])
    print("==", b['name'][:70], "total samples", tot)
    byfile = collections.Counter()
    for (f, l), n in per.items(): byfile[f] += n
    print("  by file:", [(f, round(100 * n / tot, 1)) for f, n in byfile.most_common(8)])
    allst = collections.Counter()
    for ln in stalls: allst.update(stalls[ln])
    print("  stalls:", [(k, round(100 * v / tot, 1)) for k, v in allst.most_common(8)])
    for ln, n in per.most_common(topn):
        st = ", ".join("%s %d" % (k[6:], v) for k, v in stalls[ln].most_common(3))
        print("  %5.1f%%  %-22s inst %-9d %s" % (100 * n / tot, "%s:%s" % ln, inst[ln], st))
