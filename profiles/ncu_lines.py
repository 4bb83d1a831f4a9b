"""Joins an .ncu-rep SASS page with nvdisasm line info of the built .so: executed warp instructions per source line.
usage: python profiles/ncu_lines.py file.ncu-rep lib.so kernel_substring [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# collect (file:line, inline-chain) per instruction for the section of the kernel
lines = []; cur = None; active = False
for l in dis:
    if l.startswith("//--------------------- .text."):
        active = kern in l
        continue
    if l.startswith("//---------------------") and active and ".text." not in l:
        active = False
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ie, isamp, ia = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
print("sass in report: %d, sass with line info: %d" % (len(data), len(lines)))
n = min(len(data), len(lines))
by = collections.Counter(); smp = collections.Counter(); tot = 0
for i in range(n):
    e = int(data[i][ie]); tot += e
    by[lines[i]] += e; smp[lines[i]] += int(data[i][isamp])
srcs = {}
def text(key):
    if key is None: return ""
    f, ln = key.rsplit(":", 1)
    for d in ("uu-infogr-raytracer_b200/csrc", "."):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            L = srcs[p]; k = int(ln) - 1
            return L[k].strip()[:110] if 0 <= k < len(L) else ""
    return ""
for key, e in by.most_common(top):
    print("%5.1f%% %11d smp %5d  %-22s %s" % (100.0 * e / tot, e, smp[key], key, text(key)))
