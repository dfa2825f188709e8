"""Per-source-line view of an ncu capture without the GUI: joins the SASS rows of `ncu --page source --csv` (samples and
executed-instruction counts per SASS instruction, in program order) with the line table of the same kernel from
`nvdisasm -g` (needs the object built with -lineinfo), and prints the source lines that collect the most stall samples.
  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <object.o> <mangled-name substring> [top N]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kre, obj, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may match: keep the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = blocks[int(os.environ.get("BLOCK", 0))]
hdr = blk["rows"][0]; body = blk["rows"][1:]
i_s, i_ex, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
# the section of the wanted instantiation
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and sub in l and l.rstrip().endswith(":")][0]
lines, curline = [], None
for l in sass[start + 1:]:
    if l.startswith("//--------------------- .text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curline = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,6}\*/\s+\S", l): lines.append(curline)
print(f"# {blk['name'][:90]}: {len(body)} SASS rows in the report, {len(lines)} in the object")
n = min(len(body), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_s = tot_e = 0
for k in range(n):
    s, e = int(body[k][i_s]), int(body[k][i_ex])
    a = agg[lines[k]]; a[0] += s; a[1] += e; a[2] += 1
    tot_s += s; tot_e += e
print(f"# total samples {tot_s}, warp instructions {tot_e}")
src_cache = {}
def src(fl):
    if fl is None: return ""
    f, ln = fl
    for d in (os.environ.get("SRCDIR", "phoregen_b200/csrc"), "include"):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().split("\n")
            return src_cache[p][ln - 1].strip()[:110]
    return ""
print("| samples | % | warp instr | SASS | line | source |\n|---:|---:|---:|---:|---|---|")
for fl, (s, e, c) in sorted(agg.items(), key=lambda x: -x[1][int(os.environ.get("SORT", 0))])[:top]:
    print(f"| {s} | {100.0 * s / max(tot_s, 1):.1f} | {e} | {c} | {fl[0] + ':' + str(fl[1]) if fl else '?'} | `{src(fl)}` |")
# optional: SASS context of one source line (env LINE=file:line)
if os.environ.get("LINE"):
    f, ln = os.environ["LINE"].split(":"); ln = int(ln)
    for k in range(n):
        if lines[k] == (f, ln) and int(body[k][i_s]) > tot_s * 0.002:
            print("----")
            for j in range(max(0, k - 8), min(n, k + 3)):
                print(f"{'>>' if j == k else '  '} {body[j][i_s]:>6} {body[j][i_ex]:>9} {str(lines[j][1]) if lines[j] else '?':>5} {body[j][i_src].strip()[:100]}")
