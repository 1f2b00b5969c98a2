"""Turns the ncu CSV logs brought back from the GPU box into the summaries kept under profiles/ (runs here, no GPU).

    python tools/summarize_ncu.py conv <metrics csv of tools/ncu_forward.py> <out prefix>
        per-launch table of one forward at the bench batch + profiles/conv_tc_traffic.json (what bench.py's roofline
        reads: mean DRAM bytes per launch, time-weighted tensor-pipe activity)
    python tools/summarize_ncu.py launches <gpu__time_duration csv of bench.py> <out md>
        kernel-name -> launches, total time, share of the run
"""
import csv
import io
import json
import re
import sys
from collections import OrderedDict, defaultdict


def read_ncu_csv(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.reader(io.StringIO(text[start:])))
    head = rows[0]
    return [dict(zip(head, r)) for r in rows[1:] if len(r) == len(head)]


def to_float(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def short(name):
    m = re.search(r"conv3x3_tc_kernel<([^>]*)>", name)
    if m:
        return "conv3x3_tc_kernel<" + m.group(1).replace("(bool)", "").replace("(int)", "") + ">"
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("msr::", "")[:70]


def conv(path, prefix):
    recs = read_ncu_csv(path)
    launches = OrderedDict()
    for r in recs:
        d = launches.setdefault(r["ID"], {"name": short(r["Kernel Name"])})
        val, unit = to_float(r["Metric Value"]), r["Metric Unit"]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12,
                 "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9,
                 "second": 1.0}.get(unit, 1.0)
        d[r["Metric Name"]] = val * scale
    tot_t = sum(d["gpu__time_duration.sum"] for d in launches.values())
    tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
    if not any(tp in d for d in launches.values()):   # the triage variant (part of --set full only)
        tp = "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"
    lines = ["| # | kernel | ms | tensor pipe active % | DRAM read MB | DRAM write MB | L2 bytes MB | DRAM % | SM % |", "|---|---|---|---|---|---|---|---|---|"]
    dram = 0.0
    for k, d in enumerate(launches.values()):
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        dram += rd + wr
        lines.append("| %d | `%s` | %.4f | %.1f | %.0f | %.0f | %.0f | %.1f | %.1f |" % (
            k, d["name"], d["gpu__time_duration.sum"] * 1e3, d.get(tp, 0.0), rd / 1e6, wr / 1e6,
            d.get("lts__t_bytes.sum", 0.0) / 1e6, d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0),
            d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)))
    weighted = sum(d.get(tp, 0.0) * d["gpu__time_duration.sum"] for d in launches.values()) / tot_t
    n = len(launches)
    summary = {"source": "ncu --metrics ... -k regex:'conv3x3_tc|mask_conv_tc|phase_stencil_tc' -s 53 -c 53 python "
                         "tools/ncu_forward.py 512 16 8 (second GauGAN-512 forward over 128 patches = one bench launch "
                         "group; kernels replayed alone, cold L2, boost clocks)",
               "launches": n, "patches_per_forward": 128, "total_ms_under_ncu": tot_t * 1e3,
               "dram_bytes_per_forward": dram, "dram_bytes_per_launch": dram / n,
               "tensor_pipe_active_pct_time_weighted": weighted}
    open(prefix + ".md", "w").write(
        "# tcgen05 launches of one GauGAN-512 forward at the bench batch (128 patches), ncu metrics\n\n"
        "%s\n\ntime-weighted tensor-pipe activity %.1f %%, DRAM traffic %.1f GB per forward (%.1f MB per launch), "
        "%d launches, %.2f ms under ncu (serialised, cold cache, boost clocks: compare shares, not absolutes)\n\n" %
        (summary["source"], weighted, dram / 1e9, dram / n / 1e6, n, tot_t * 1e3) + "\n".join(lines) + "\n")
    big = [d for d in launches.values() if d["gpu__time_duration.sum"] > 0.3e-3]
    summary["tensor_pipe_active_pct_time_weighted_launches_over_0.3ms"] = (
        sum(d.get(tp, 0.0) * d["gpu__time_duration.sum"] for d in big) / max(sum(d["gpu__time_duration.sum"] for d in big), 1e-12))
    json.dump(summary, open("profiles/conv_tc_traffic.json", "w"), indent=1)
    print(json.dumps(summary, indent=1))


def launches(path, out):
    recs = read_ncu_csv(path)
    fam = defaultdict(lambda: [0, 0.0])
    for r in recs:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        unit = r["Metric Unit"]
        scale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9,
                 "second": 1.0}.get(unit, 1.0)
        f = fam[short(r["Kernel Name"])]
        f[0] += 1
        f[1] += to_float(r["Metric Value"]) * scale
    tot = sum(v[1] for v in fam.values())
    rows = sorted(fam.items(), key=lambda kv: -kv[1][1])
    text = ["# ncu launch list of bench.py (gpu__time_duration.sum, --clock-control none): per kernel", "",
            "total %d launches, %.1f ms of kernel time (serialised, cold cache)" % (sum(v[0] for v in fam.values()), tot * 1e3),
            "", "| kernel | launches | ms | share |", "|---|---|---|---|"]
    for name, (cnt, t) in rows:
        text.append("| `%s` | %d | %.2f | %.2f %% |" % (name, cnt, t * 1e3, 100 * t / tot))
    tc = sum(t for name, (cnt, t) in rows if any(k in name for k in ("conv3x3_tc", "mask_conv_tc", "phase_stencil_tc")))
    text += ["", "tcgen05 convolution family: %.2f %% of the kernel time" % (100 * tc / tot)]
    open(out, "w").write("\n".join(text) + "\n")
    print("\n".join(text[:14]))


if __name__ == "__main__":
    {"conv": conv, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
