"""Summarise gpurun_out/*.ncu-rep and the launch list into profiles/ (run in the build
container: ncu reads reports without a GPU).

    python tools/ncu_summary.py <round tag> <launches.csv> <name=report.ncu-rep> ...
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def source_page(rep, top=14):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        return []
    hdr, data = rows[hi[0]], [r for r in rows[hi[0] + 1:] if len(r) > 5]
    ia, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
    cols = {k: hdr.index(k) for k in ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_math",
                                      "stall_mio", "stall_lg") if k in hdr}
    tot = sum(int(r[ia] or 0) for r in data) or 1
    out = []
    for r in sorted(data, key=lambda r: -int(r[ia] or 0))[:top]:
        why = max(cols, key=lambda k: int(r[cols[k]] or 0))
        out.append((100.0 * int(r[ia] or 0) / tot, why, r[isrc].strip()[:100]))
    return out


def main():
    tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(PROF, exist_ok=True)
    lines = [f"# ncu summary, round {tag}", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` launch list (cold cache, serialised: compare "
             "shares, not absolutes) and `ncu --set full --clock-control none --import-source on` captures of the two "
             "dominant kernels; produced by `tools/ncu_summary.py` from the reports brought back in `gpurun_out/`.", ""]
    if os.path.exists(launches):
        rows = list(csv.reader(open(launches)))
        hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
        hdr = rows[hi]
        kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in rows[hi + 1:]:
            if len(r) <= mv:
                continue
            v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
            agg.setdefault(r[kn].split("(")[0].replace("void ", "").replace("femb::", "")[:58], []).append(v)
        tot = sum(sum(v) for v in agg.values())
        lines += ["## Launch list (`bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline --e2e-steps 1`, first 420 launches)", "",
                  "| kernel | launches | mean µs | share of GPU time |", "|---|---:|---:|---:|"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {100*sum(v)/tot:.1f} % |")
        lines.append("")
    traffic = {}
    for item in reps:
        name, rep = item.split("=")
        recs, units = raw_page(rep)
        if not recs:
            continue
        d = recs[0]
        lines += [f"## `{name}`: {d.get('Kernel Name', '')[:110]}", "", "| metric | value |", "|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| `{k}` | {d[k]} {units.get(k, '')} |")
        rd = float(d["dram__bytes_read.sum"].replace(",", "")) * TO_BYTES[units["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"].replace(",", "")) * TO_BYTES[units["dram__bytes_write.sum"]]
        traffic[name] = rd + wr
        lines += [f"| **DRAM traffic per launch** | {(rd+wr)/1e9:.3f} GB (read {rd/1e9:.3f} + write {wr/1e9:.3f}) |", ""]
        stalls = {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(v) for k, v in d.items()
                  if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        lines += ["Warp stalls per issued instruction: " + ", ".join(f"{k} {v:.2f}" for k, v in top), ""]
        src = source_page(rep)
        if src:
            lines += ["Top sampled instructions (share of samples, dominant stall, SASS):", "", "```"]
            lines += [f"{p:5.1f}%  {why:<15s} {s}" for p, why, s in src] + ["```", ""]
    with open(os.path.join(PROF, f"{tag}_summary.md"), "w") as f:
        f.write("\n".join(lines))
    # profiles/traffic.json: {"<n>": {"assemble": bytes, "spmv": bytes}} -- names of the form <kernel>@<n> are recorded
    tj = os.path.join(PROF, "traffic.json")
    old = json.load(open(tj)) if os.path.exists(tj) else {}
    old = {k: v for k, v in old.items() if isinstance(v, dict)}
    for name, b in traffic.items():
        if "@" in name:
            kern, n = name.split("@")
            old.setdefault(n, {})[kern] = b
    json.dump(old, open(tj, "w"), indent=1)
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
