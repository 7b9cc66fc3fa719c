"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_bench.csv  > profiles/rN_launches_bench.md
  python tools/ncu_summary.py full     gpurun_out/lin_full.ncu-rep    > profiles/rN_lin_full.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("stba::", "")
    return name.strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        k = short(r[ik])
        ns = float(r[iv].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += ns
        total += ns
    print("| kernel | launches | total us | mean us | share | grid (last) | block |")
    print("|---|---|---|---|---|---|---|")
    for k, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f %% | %s | %s |" % (k, n, ns / 1e3, ns / n / 1e3, 100 * ns / total, g, b))
    print("\n%d launches, %.3f ms of serialised kernel time (cold-cache: compare shares, not absolutes)" % (len(rows) - 1, total / 1e6))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print("### `%s`  (launch id %s)\n" % (short(d[idx["Kernel Name"]]), d[idx["ID"]]))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in idx and d[idx[k]] != "":
                print("| %s | %s | %s |" % (k, d[idx[k]], units[idx[k]]))
        st = []
        for h, i in idx.items():
            m = STALL.match(h)
            if m and "not_issued" not in h and d[i] not in ("", "0"):
                try:
                    st.append((float(d[i].replace(",", "")), m.group(1) or m.group(2)))
                except ValueError:
                    pass
        st.sort(reverse=True)
        if st:
            print("\nwarp stall reasons (cycles per issued instruction): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
        print()


def traffic(path):
    """profiles/lin_traffic.json: DRAM bytes per launch of lin_lm2 + lin_cam2 (mean over the captured launches)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for d in data:
        k = short(d[idx["Kernel Name"]])
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[idx[m]].replace(",", "")) * scale[units[idx[m]]]
        per.setdefault(k, []).append(tot)
    kern = {k: sum(v) / len(v) for k, v in per.items()}
    lin3 = [v for k, v in kern.items() if k.startswith("k_lin3")]
    if lin3:
        lin, what = sum(lin3), "k_lin3"
    else:
        lin, what = sum(v for k, v in kern.items() if k.startswith("k_lin_lm2<0") or k.startswith("k_lin_cam2")), "k_lin_lm2 + k_lin_cam2"
    print(json.dumps({"bytes_per_linearisation": lin, "per_kernel": kern, "source": "ncu --set full capture " + path.split("/")[-1] +
                      " (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch of " + what + ")"}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
