"""Turn ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r02_launches.csv profiles/r02_launches_step.txt [first_id last_id]
    python tools/summarize_ncu.py full gpurun_out/r02_conv.ncu-rep profiles/r02_ncu_conv.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__warps_active.avg.per_cycle_active", "warps active / scheduler"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem"),
]


def short(name):
    n = name.split("(")[0]
    for p in ("void ", "scn::", "tma::", "at::native::", "cub::CUB_200802_SM_1000::"):
        n = n.replace(p, "")
    return n[:60]


def launches(src, dst, first=None, last=None):
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("nsecond", "ns") else v / 1e3 if unit in ("usecond", "us") else v
        rows.append((int(r["ID"]), r["Kernel Name"], ms))
    if first is not None:
        rows = [r for r in rows if first <= r[0] <= last]
    agg = defaultdict(lambda: [0.0, 0])
    for _, k, ms in rows:
        agg[short(k)][0] += ms
        agg[short(k)][1] += 1
    total = sum(v[0] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"launches {len(rows)}, total {total:.2f} ms (ids {rows[0][0]}..{rows[-1][0]}); ncu --metrics gpu__time_duration.sum, "
                f"cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{ms:9.3f} ms {100 * ms / total:5.1f}%  x{n:<4d} {k}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"ncu --set full --clock-control none ({src}); one block per captured launch\n")
        for r in rows[2:]:
            f.write(f"\n== {short(r[hdr.index('Kernel Name')])}  (launch id {r[hdr.index('ID')]})\n")
            for k, label in KEYS:
                if k in hdr:
                    f.write(f"   {label:28s} {r[hdr.index(k)]} {units[hdr.index(k)]}\n")
            stalls = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr)
                             if "issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)[:5]
            f.write("   top stall reasons (warps per issue): " +
                    ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in stalls) + "\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], *(int(a) for a in sys.argv[4:6]))
    else:
        full(sys.argv[2], sys.argv[3])
