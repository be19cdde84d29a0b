"""Ranks source lines of a kernel by executed instructions / stall samples:
joins an ncu --page source CSV (SASS view) with nvdisasm -g line info of the built object."""
import csv, collections, re, subprocess, sys, os, tempfile, glob
rep, obj, pat, srcfile = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(tmp + "/*.cubin"), capture_output=True, text=True).stdout
cur = None; amap = {}; infn = False
for ln in dis.splitlines():
    if ln.strip().startswith(".text."): infn = pat in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infn: amap[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ie = hdr.index("Instructions Executed"); smp = hdr.index("# Samples"); ad = hdr.index("Address")
by = collections.Counter(); bys = collections.Counter(); tot = tots = 0; base = None
for r in rows[2:]:
    try: n = int(r[ie] or 0); s = int(r[smp] or 0); a = int(r[ad], 16)
    except Exception: continue
    if base is None: base = a
    key = amap.get(a - base, ("?", 0)); by[key] += n; bys[key] += s; tot += n; tots += s
src = open(srcfile).read().split("\n")
print(f"total warp instructions {tot}  samples {tots}")
for (f, l), n in by.most_common(top):
    text = src[l - 1].strip()[:100] if f == os.path.basename(srcfile) and 0 < l <= len(src) else f
    print(f"{100*n/tot:5.1f}% inst {100*bys[(f,l)]/max(tots,1):5.1f}% smp  L{l:<4d} {text}")
