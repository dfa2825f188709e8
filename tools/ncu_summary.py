"""Summarise ncu outputs into small text files under profiles/ (tracked).
  python tools/ncu_summary.py launches <launches.csv> <out.md> [--skip N]
  python tools/ncu_summary.py report <file.ncu-rep> <out.md>
"""
import collections, csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^.*::", "", name)
    return re.sub(r"<.*", "", name)


def launches(path, out, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= iv or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip:
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1e-6)
        k = short(r[ik]) + ("<" + re.sub(r"^[^<]*<", "", r[ik]).split(">")[0] + ">" if "<" in r[ik] else "")
        t = tot.setdefault(k, [0.0, 0])
        t[0] += v; t[1] += 1
    total = sum(t[0] for t in tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path}; gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)\n\n")
        f.write(f"launches counted: {sum(t[1] for t in tot.values())} (first {skip} skipped), total {total:.3f} ms\n\n| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, (ms, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| {k} | {c} | {ms:.3f} | {ms / c:.4f} | {100 * ms / total:.1f}% |\n")
    print(open(out).read())


def report(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {path}\n\n")
        for r in rows[2:]:
            f.write(f"## {short(r[hdr.index('Kernel Name')])}  (launch id {r[0]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            for i, h in enumerate(hdr):
                if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                    try:
                        if float(r[i]) >= 3.0:
                            f.write(f"| {h} | {r[i]} | % |\n")
                    except ValueError:
                        pass
            f.write("\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
        launches(sys.argv[2], sys.argv[3], skip)
    else:
        report(sys.argv[2], sys.argv[3])
