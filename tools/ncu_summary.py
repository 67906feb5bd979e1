#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [-o profiles/name.txt]

Per profiled launch: duration, launch geometry, issue / pipe utilisation, DRAM + L2 traffic, hit rates; for the
first launch of each kernel: SASS opcode mix and warp-stall distribution from the source page."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"), ("sm__cycles_active.avg", "SM cycles active (avg)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe fp64 %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe xu %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "pipe tensor %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    out = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    lines = []
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines.append("ncu report: %s  (%d profiled launches; --set full --clock-control none)" % (rep, len(rows)))
    for li, r in enumerate(rows):
        lines.append("")
        lines.append("launch %d: %s" % (li, r[ix["Kernel Name"]][:110]))
        for k, label in KEYS:
            if k in ix:
                lines.append("  %-32s %14s %s" % (label, r[ix[k]], units[ix[k]]))
    seen = set()
    for li, r in enumerate(rows):
        name = r[ix["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"]))))
        hi = [i for i, rr in enumerate(src) if any("Instructions Executed" == c for c in rr)]
        if not hi:
            continue
        h = src[hi[0]]
        sx = {n: i for i, n in enumerate(h)}
        ops, stalls = collections.Counter(), collections.Counter()
        stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        seen_addr = set()
        for rr in src[hi[0] + 1:]:
            if len(rr) < len(h) or rr[sx["Address"]] in seen_addr:
                continue
            seen_addr.add(rr[sx["Address"]])
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", rr[sx["Source"]].strip())
            if not m:
                continue
            try:
                n = int(rr[sx["Instructions Executed"]])
            except ValueError:
                continue
            ops[".".join(m.group(2).split(".")[:3])] += n
            for c in stall_cols:
                try:
                    stalls[c] += int(rr[sx[c]])
                except ValueError:
                    pass
        tot = sum(ops.values()) or 1
        lines.append("")
        lines.append("SASS opcode mix of launch %d (%s), %d warp instructions:" % (li, name[:60], tot))
        for k, v in ops.most_common(16):
            lines.append("  %-24s %10d %5.1f%%" % (k, v, 100.0 * v / tot))
        st = sum(stalls.values()) or 1
        lines.append("warp-stall samples: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / st) for k, v in stalls.most_common(8)))
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
