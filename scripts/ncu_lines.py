"""Attribute an .ncu-rep's per-instruction samples / instruction counts to CUDA source lines.
Usage: ncu_lines.py rep kernel-regex mangled-substring [cubin]   (needs -lineinfo; joins ncu's SASS page with nvdisasm -g by instruction index)"""
import csv, io, re, subprocess, sys, collections, glob, os, tempfile
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diff-dope_b200", "diffdope", "_lib", "libddope_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if (".text." in out) and (mangled in out):
        sec = out.split("\n")
        start = next(i for i, l in enumerate(sec) if l.startswith(".text.") and mangled in l)
        cur = None; lines = []
        for l in sec[start + 1:]:
            if l.startswith("//---") or l.startswith(".text."): break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
        break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ie, sm = h.index("Instructions Executed"), h.index("# Samples")
data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
print("sass instructions: ncu", len(data), "nvdisasm", len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) and lines[i] else ("?", 0)
    agg[key][0] += int(r[ie]); agg[key][1] += int(r[sm])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values()) or 1
srcs = {}
top = sorted(agg.items(), key=lambda kv: -kv[1][0 if os.environ.get("BY_INST") else 1])[: int(sys.argv[4]) if len(sys.argv) > 4 else 40]
for (f, ln), (ins, sa) in top:
    if f not in srcs:
        pth = os.path.join(os.path.dirname(lib), "..", "..", "csrc", f)
        srcs[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
    text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * sa / ts, 100 * ins / ti, f, ln, text))
