#!/usr/bin/env python
"""Turns an ncu report (or a launch-list CSV) into the small text summaries kept under profiles/.

    python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep   > profiles/<name>.md
    python tools/ncu_summary.py list  gpurun_out/launches.csv   > profiles/<name>.md

Runs in the build container (ncu is installed, no GPU needed to read a report).
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs (32 B each)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instructions"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stall lg_throttle %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall barrier %"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full summary of `%s`\n" % path)
    print("Captured with `--clock-control none`; per-launch values (cold-cache, serialised replays).\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("## %s\n" % name)
        print("| metric | value |\n|---|---|")
        for key, label in FULL_METRICS:
            if key in hdr:
                i = hdr.index(key)
                print("| %s (`%s`) | %s %s |" % (label, key, r[i], units[i]))
        print()


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        k = r[ki].split("(")[0]
        agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("# kernel launch list of `%s`\n" % path)
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` over a bench.py run; times are "
          "cold-cache and serialised -- compare SHARES.\n")
    print("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %d | %.3f | %.4f | %.1f %% |" % (k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6, 100 * sum(v) / tot))


if __name__ == "__main__":
    (full if sys.argv[1] == "full" else launch_list)(sys.argv[2])
