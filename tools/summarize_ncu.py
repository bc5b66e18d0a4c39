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


def short(name):
    name = name.replace("void ", "").replace("acx::", "")
    return name.split("(")[0][:70]


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    cnt = collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v          # -> us
        k = short(r[ik])
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
