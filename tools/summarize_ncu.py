"""Turn ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md>      # per-kernel time shares of one forward
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md>         # key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def is_torch(name):
    return any(t in name for t in ("at::", "at_cuda", "cub::", "elementwise", "vectorized", "CatArray", "memcpy"))


def short(name):
    name = name.replace("void ", "").replace("acx::", "")
    return name.split("(")[0][:70]


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    cnt = collections.Counter()
    other = 0.0
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v          # -> us
        if is_torch(r[ik]):
            other += v                                            # torch plumbing (input synthesis, output copies)
            continue
        k = short(r[ik])
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
    if other:
        tot["(torch plumbing: weight repack, input synthesis, output copies)"] = other
        cnt["(torch plumbing: weight repack, input synthesis, output copies)"] = 0
    total = sum(tot.values())
    with open(out, "w") as fh:
        fh.write(f"ncu launch list `{path}` ({sum(cnt.values())} launches, cold-cache / serialised: compare SHARES)\n\n")
        fh.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            fh.write(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% |\n")
        fh.write(f"| **all** | {sum(cnt.values())} | {total:.1f} | 100% |\n")
    print(open(out).read())


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    with open(out, "w") as fh:
        fh.write(f"`ncu --set full --clock-control none` capture `{path}` (per launch)\n\n")
        for r in rows[2:]:
            fh.write(f"### `{short(r[ik])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    fh.write(f"| {label} (`{key}`) | {r[i]} | {units[i]} |\n")
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")])
                wr = float(r[hdr.index("dram__bytes_write.sum")])
                u = units[hdr.index("dram__bytes_read.sum")]
                fh.write(f"| **traffic = dram read + write** | {rd + wr:.2f} | {u} |\n")
            except (ValueError, KeyError):
                pass
            fh.write("\n")
    print(open(out).read()[:3000])


def rawcsv(path, out):
    """`ncu -i rep --page raw --csv` output (made on the GPU box: the .ncu-rep of a whole forward is > 64 MiB):
    per kernel NAME, the launch count and launch-averaged key metrics."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    groups = collections.OrderedDict()
    for r in rows[2:]:
        groups.setdefault(short(r[ik]), []).append(r)
    def col(rs, key):
        i = hdr.index(key)
        vals = []
        for r in rs:
            try:
                vals.append(float(r[i].replace(",", "")))
            except ValueError:
                pass
        return (sum(vals) / len(vals) if vals else float("nan")), units[i]
    tot_time = sum(col(rs, "gpu__time_duration.sum")[0] * len(rs) for rs in groups.values())
    with open(out, "w") as fh:
        fh.write(f"`ncu --set full --clock-control none` over ONE forward of 64 clips (all {len(rows) - 2} launches, "
                 f"CUDA-graph replay off); launch-averaged per kernel. Source: `{path}`\n\n")
        fh.write("| kernel | n | avg dur | share | dram rd+wr / launch | dram % | tensor % | fma % | issue % | occ % | regs |\n")
        fh.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, rs in sorted(groups.items(), key=lambda kv: -col(kv[1], "gpu__time_duration.sum")[0] * len(kv[1])):
            if is_torch(rs[0][ik]):
                continue
            d, du = col(rs, "gpu__time_duration.sum")
            rd, ru = col(rs, "dram__bytes_read.sum")
            wr, _ = col(rs, "dram__bytes_write.sum")
            fh.write(f"| `{k}` | {len(rs)} | {d:.1f} {du} | {100 * d * len(rs) / tot_time:.1f}% | {rd + wr:.1f} {ru} | "
                     f"{col(rs, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[0]:.1f} | "
                     f"{col(rs, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[0]:.1f} | "
                     f"{col(rs, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')[0]:.1f} | "
                     f"{col(rs, 'smsp__issue_active.avg.pct_of_peak_sustained_active')[0]:.1f} | "
                     f"{col(rs, 'sm__warps_active.avg.pct_of_peak_sustained_active')[0]:.1f} | "
                     f"{col(rs, 'launch__registers_per_thread')[0]:.0f} |\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "rawcsv": rawcsv}[sys.argv[1]](sys.argv[2], sys.argv[3])
