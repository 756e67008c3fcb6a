"""Correlate an ncu report's per-SASS counters with source lines (needs the same .so that ran)."""
import collections, csv, io, re, subprocess, sys
rep, cubin_name, kernel_re = sys.argv[1], sys.argv[2], sys.argv[3]
sel = sys.argv[4] if len(sys.argv) > 4 else None
subprocess.run("mkdir -p /tmp/cub && cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all /root/repo/simple_pose_b200/lib/libsimple_pose_b200.so > /dev/null", shell=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", "/tmp/cub/" + cubin_name], capture_output=True, text=True).stdout
cur_fn = cur_line = None
line_of = {}
for ln in sass.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur_line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and cur_fn: line_of[(cur_fn, int(m.group(1), 16))] = (cur_line, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_re], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')
for blk in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    kname = rows[0][1]
    if sel and sel not in kname: continue
    hdr = rows[1]
    ia, ie, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    # mangled name lookup: pick fn whose line map has most addresses matching
    fns = set(k[0] for k in line_of)
    def demangle_match(f):
        parts = re.findall(r'[A-Za-z_]+', kname.split('(')[0])
        return all(p in f for p in parts[-1:])
    cands = [f for f in fns if demangle_match(f)]
    # choose by template args
    targs = re.findall(r'\(bool\)(\d)', kname) + re.findall(r'\(int\)(\d+)', kname)
    def score(f):
        return sum(1 for a in targs if ('Lb%sE' % a in f) or ('Li%sE' % a in f))
    tag = ''.join('Lb%sE' % a for a in re.findall(r'\(bool\)(\d)', kname))
    cands2 = [f for f in cands if tag in f] or cands
    fn = cands2[0]
    agg = collections.Counter(); samp = collections.Counter(); stalls = collections.defaultdict(collections.Counter)
    base = None; tot = 0; totsamp = 0
    for r in rows[2:]:
        try: addr = int(r[ia], 16)
        except Exception: continue
        if base is None: base = addr
        n = int(r[ie] or 0); s = int(r[isamp] or 0)
        info = line_of.get((fn, addr - base))
        key = info[0] if info and info[0] else ('?', 0)
        agg[key] += n; samp[key] += s; tot += n; totsamp += s
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: stalls[key][hdr[i]] += v
    print("==", kname[:100], "inst", tot, "samples", totsamp)
    allst = collections.Counter()
    for k in stalls: allst.update(stalls[k])
    print("stall totals:", [(k, v) for k, v in allst.most_common(8)])
    for k, v in sorted(agg.items(), key=lambda kv: (-kv[1] if "BYINST" in __import__("os").environ else -samp[kv[0]]))[:int(__import__("os").environ.get("TOPN","22"))]:
        top = stalls[k].most_common(2)
        print("  %-28s inst %5.1f%%  samples %5.1f%%  %s" % ("%s:%d" % k, 100 * v / max(1, tot), 100 * samp[k] / max(1, totsamp), top))
