#!/usr/bin/env python
"""Map an ncu SASS source page (--page source --csv) onto CUDA source lines using nvdisasm -g line
markers of the same cubin.  Usage: tools_ncu_lines.py <rep> <kernel-substring> [top]"""
import csv, re, subprocess, sys, os, tempfile, collections
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(os.environ.get("XZB200_LIB", "xz_b200/libxzb200.so"))], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = max([f for f in os.listdir(tmp) if f.endswith(".cubin")], key=lambda f: os.path.getsize(os.path.join(tmp, f)))
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.splitlines()
addr2line = {}
infn = False; cur = None
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        infn = kname in ln
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m: addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows[:5]) if "Address" in r][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ti = ts = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    if base is None: base = a
    key = addr2line.get(a - base)
    inst = float(r[ci["Instructions Executed"]] or 0); samp = float(r[ci["# Samples"]] or 0)
    agg[key][0] += inst; agg[key][1] += samp; ti += inst; ts += samp
    for s in stall_cols:
        v = float(r[ci[s]] or 0)
        if v: agg[key][2][s] += v
print(f"total warp-instructions {ti:.0f}, samples {ts:.0f}")
src_cache = {}
def src(key):
    if not key: return ""
    f, l = key
    for d in ("xz_b200/csrc", "."):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if l - 1 < len(src_cache[p]) else ""
    return ""
for key, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    s3 = ",".join(f"{k[6:]}:{int(v)}" for k, v in st.most_common(3))
    print(f"{100*samp/ts:5.1f}%s {100*inst/ti:5.1f}%i {key} [{s3}] {src(key)}")

# ---- optional: aggregate by enclosing function (crude: last "__device__" line above) ----
if len(sys.argv) > 4 and sys.argv[4] == "func":
    import bisect
    fn_starts = {}
    def fn_of(key):
        if not key: return "?"
        f, l = key
        for d in ("xz_b200/csrc", "."):
            p = os.path.join(d, f)
            if os.path.exists(p):
                if p not in fn_starts:
                    lines = open(p).read().splitlines()
                    st = [(i + 1, re.sub(r"\s+", " ", ln.strip())[:70]) for i, ln in enumerate(lines) if re.match(r"\s*(static )?__device__|^XZB_HD|^__global__|\s*XZB_HDM", ln)]
                    fn_starts[p] = st
                st = fn_starts[p]
                idx = bisect.bisect_right([s[0] for s in st], l) - 1
                return f"{f}:{st[idx][1]}" if idx >= 0 else f
        return f
    fagg = collections.defaultdict(lambda: [0.0, 0.0])
    for key, (inst, samp, st) in agg.items():
        k = fn_of(key); fagg[k][0] += inst; fagg[k][1] += samp
    print("---- by function ----")
    for k, (inst, samp) in sorted(fagg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{100*samp/ts:5.1f}%s {100*inst/ti:5.1f}%i  {k}")
