"""Condense ncu outputs into the small text summaries committed under profiles/.

  python tools/summarise_ncu.py launches <launches.csv> <out.md>     # per-kernel shares of a launch list
  python tools/summarise_ncu.py full <file.ncu-rep> <out.md>         # key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = name.replace("void ", "").replace("mamg::", "")
    return name.split("(")[0]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        try:
            t = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = short(r[ik])
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# ncu launch list summary of `{path}`\n\n")
        fh.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n\n")
        fh.write(f"launches: {sum(v[0] for v in agg.values())}, total kernel time {tot / 1e6:.3f} ms\n\n")
        fh.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {c} | {t / 1e6:.3f} | {t / tot:.3f} | {t / c / 1e3:.1f} |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary of `{path}`\n\n")
        for r in rows[2:]:
            fh.write(f"## `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write(f"| {k} | {r[i]} | {units[i]} |\n")
            fh.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
