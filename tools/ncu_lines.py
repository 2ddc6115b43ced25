"""Attribute an ncu --import-source report to source lines offline.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel base name> <launch index among that kernel's launches> [top N] [stall column]
       [range:<file>:<lo>-<hi> ...]
Joins `ncu --page source --csv` (per-SASS 'Instructions Executed', stall samples) with `nvdisasm -g` line info of the cubin inside
libsdfmesh.so (must be the same build as the profiled one).  A range sums every instruction whose address lies between the
first and the last instruction attributed to those lines (captures inlined callees inside a loop body)."""
import collections, csv, io, os, pathlib, re, subprocess, sys, tempfile

rep, kern, skip = sys.argv[1], sys.argv[2], int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
stall_col = sys.argv[5] if len(sys.argv) > 5 and not sys.argv[5].startswith("range:") else "# Samples"
so = os.environ.get("SDM_LIB", "bevy-signed-distance-mesh-generation_b200/libsdfmesh.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = next(pathlib.Path(tmp).glob("*.cubin"))
dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
line_of = {}
cur = None; infn = False
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        infn = re.search(r"\d+" + kern + "E", l) is not None
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (pathlib.Path(m.group(1)).name, int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function", "--kernel-name", kern,
                      "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index(stall_col)
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(data[0][ia], 16)
by_line = collections.Counter(); samp = collections.Counter()
tot = 0; tots = 0
for r in data:
    off = int(r[ia], 16) - base
    n = int(r[ii] or 0); s = int(r[isamp] or 0)
    k = line_of.get(off)
    by_line[k] += n; samp[k] += s; tot += n; tots += s
print(f"{kern}[{skip}]: total warp instructions {tot:,}  {stall_col} {tots:,}")
src_cache = {}
def src(k):
    if not k: return ""
    f = pathlib.Path("bevy-signed-distance-mesh-generation_b200/csrc") / k[0]
    if f not in src_cache: src_cache[f] = f.read_text().splitlines() if f.exists() else []
    L = src_cache[f]
    return L[k[1] - 1].strip()[:100] if 0 < k[1] <= len(L) else ""
order = samp if stall_col != "# Samples" else by_line
for k, _ in order.most_common(top):
    print(f"{by_line[k] / tot * 100:5.1f}% inst {samp[k] / max(tots, 1) * 100:5.1f}% {stall_col[:12]}  {k[0] if k else '?'}:{k[1] if k else 0}: {src(k)}")
for spec in sys.argv[5:]:
    if not spec.startswith("range:"): continue
    _, f, lr = spec.split(":")
    lo, hi = map(int, lr.split("-"))
    offs = [o for o, k in line_of.items() if k and k[0] == f and lo <= k[1] <= hi]
    if not offs: print(spec, "no instructions"); continue
    a, b = min(offs), max(offs)
    n = sum(int(r[ii] or 0) for r in data if a <= int(r[ia], 16) - base <= b)
    s = sum(int(r[isamp] or 0) for r in data if a <= int(r[ia], 16) - base <= b)
    print(f"{spec}: addresses {a:#x}..{b:#x} ({(b - a) // 16 + 1} instrs): {n / tot * 100:.1f}% of instructions, {s / max(tots, 1) * 100:.1f}% of {stall_col}")
