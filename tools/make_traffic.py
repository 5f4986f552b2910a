"""profiles/traffic.json from ncu captures, stamped with the kernel sources it was taken from.

bench.py's `roofline` object needs, per workload, numbers only a profiler sees: warp instructions, issue-active,
DRAM / L2 / L1 bytes of the dominant kernel.  They are fixed for a workload AND A BUILD, so every entry carries the
sha256 of the kernel sources (csrc/*.cu, csrc/*.cuh, include/rtx_b200.h) it was measured on; bench.py recomputes
the hash and emits null instead of stale numbers when it differs.

usage:
    python tools/make_traffic.py <workload> <report.ncu-rep> <kernel-regex> [launches.csv [companion-regex ...]]

      workload     c3 | c4 | c5 | c3@8 ... (key in traffic.json; the entry is replaced, other entries are kept; "@N" = one
                   rank's share of an N-GPU frame, captured from tools/rank_timing.py N)
      report       ncu --set full capture of the bench/sweep command for that workload (one GPU)
      kernel-regex picks the dominant kernel's launch inside the report (first match is used)
      launches.csv optional `ncu --metrics gpu__time_duration.sum` launch list of the same command:
                   with the regexes of the other kernels of one step (collect passes, overflow launch, ray tables)
                   gives the dominant kernel's share of the step (median duration per kernel)
"""
import csv
import datetime
import glob
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_source_sha():
    """sha256 (first 16 hex digits) over the sources every kernel is compiled from."""
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "opencl_raytracer_b200", "csrc", "*.cu")) +
                   glob.glob(os.path.join(ROOT, "opencl_raytracer_b200", "csrc", "*.cuh")) +
                   [os.path.join(ROOT, "include", "rtx_b200.h")])
    for fn in files:
        h.update(os.path.basename(fn).encode())
        with open(fn, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    except OSError:
        return None


def raw_page(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(s):
    return float(s.replace(",", "")) if s not in ("", "n/a") else None


def scaled(value, unit):
    """ncu prints byte counts as Kbyte / Mbyte / Gbyte and times in ns / us / ms"""
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return value * m.get(unit, 1)


def main():
    workload, report, pattern = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
    launches = sys.argv[4] if len(sys.argv) > 4 else None
    hdr, units, data = raw_page(report)
    kn = hdr.index("Kernel Name")
    rows = [r for r in data if pattern.search(r[kn])]
    if not rows:
        sys.exit("no launch matches %r; kernels in the report: %s" % (pattern.pattern, sorted({r[kn][:70] for r in data})))
    r = rows[0]

    def get(name):
        if name not in hdr:
            return None
        i = hdr.index(name)
        v = num(r[i])
        return None if v is None else scaled(v, units[i])

    sectors_l1 = get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
    sectors_l2 = get("lts__t_sectors_srcunit_tex_op_read.sum")
    entry = {
        "stamp": {"kernel_src_sha16": kernel_source_sha(), "git_head_at_capture": git_head(),
                  "captured_utc": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ"), "report": os.path.basename(report), "n_gpus": 1},
        "kernel": r[kn],
        "duration_ms": get("gpu__time_duration.sum"),
        "warp_instructions": get("smsp__inst_executed.sum"),
        "thread_instructions": get("smsp__thread_inst_executed.sum"),
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "avg_threads_per_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": get("launch__registers_per_thread"),
        "dram_bytes_read": get("dram__bytes_read.sum"),
        "dram_bytes_write": get("dram__bytes_write.sum"),
        "l1_load_bytes": sectors_l1 * 32 if sectors_l1 is not None else None,
        "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"),
        "l1tex_throughput_pct": get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_read_bytes_from_l1": sectors_l2 * 32 if sectors_l2 is not None else None,
        "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
        "stall_long_scoreboard": get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    }
    if launches:
        with open(launches) as f:
            lines = [ln for ln in f if ln.startswith('"')]
        rd = list(csv.DictReader(io.StringIO("".join(lines))))
        per = {}
        for row in rd:
            if row.get("Metric Name") == "gpu__time_duration.sum":
                per.setdefault(row["Kernel Name"], []).append(float(row["Metric Value"].replace(",", "")))
        mine = [k for k in per if pattern.search(k)]
        if mine:
            # share of one step: medians per kernel (the list may hold other launches of the same command: warm-up,
            # end-to-end bands, other code paths), dominant kernel against the companions named on the command line
            import statistics
            ours = statistics.median(per[mine[0]])
            whole, companions = ours, {}
            for pat in sys.argv[5:]:
                ks = [k for k in per if re.search(pat, k) and k not in mine]
                if ks:
                    companions[ks[0][:60]] = statistics.median(per[ks[0]]) * 1e-6
                    whole += statistics.median(per[ks[0]])
            entry["share_of_kernel_ms_pct"] = 100.0 * ours / whole
            entry["companions_ms"] = companions
            entry["launch_list"] = os.path.basename(launches)
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            allw = json.load(f)
    except OSError:
        allw = {}
    allw["_comment"] = ("Per-launch profiler numbers of each workload's dominant kernel, written by tools/make_traffic.py from ncu --set full "
                        "captures.  Every entry is stamped with the sha256 of the kernel sources it was measured on; bench.py ignores an "
                        "entry whose stamp differs from the sources of the build it runs.")
    if "@" in workload:
        entry["stamp"]["what"] = "rank 0 of %s ranks, rendered on one GPU (tools/rank_timing.py): the launch a rank of the N-GPU run executes" % workload.split("@")[1]
    allw[workload] = entry
    with open(path, "w") as f:
        json.dump(allw, f, indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
