#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on) into the text summary kept
under profiles/: headline counters, issue-stall shares and the SASS opcode mix of the
chosen kernel.  Runs on the CPU box (`ncu -i`); no GPU needed.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--kernel engine] [--rows N]
"""
import argparse
import collections
import csv
import io
import re
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return out


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, "raw"))))
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    return rows[0], rows[1], rows[2:]


def source_tables(rep):
    """yield (kernel name, header, rows) per kernel of the source page"""
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, "source"))))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name, hdr = rows[i][1], rows[i + 1]
            j = i + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                j += 1
            yield name, hdr, rows[i + 2:j]
            i = j
        else:
            i += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", default="engine", help="substring of the kernel name")
    ap.add_argument("--rows", type=float, default=0, help="transform rows per launch (per-row instruction counts)")
    ap.add_argument("--alg-bytes", type=float, default=0, help="algorithmic bytes per launch")
    args = ap.parse_args()

    hdr, units, rows = raw_rows(args.rep)
    ki = hdr.index("Kernel Name")
    sel = [r for r in rows if args.kernel in r[ki]]
    print(f"# ncu summary of {args.rep}\n")
    for r in sel[:1]:
        print(f"kernel: {r[ki]}\n")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        t_ns = float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * \
            {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(units[hdr.index("gpu__time_duration.sum")], 1)
        rd, wr = (r[hdr.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = float(rd) * scale[units[hdr.index("dram__bytes_read.sum")]] + \
            float(wr) * scale[units[hdr.index("dram__bytes_write.sum")]]
        print(f"\n  DRAM traffic (read+write) per launch: {tr / 1e9:.3f} GB in {t_ns / 1e6:.3f} ms "
              f"(under ncu, cold) = {tr / t_ns:.1f} GB/s")
        if args.alg_bytes:
            print(f"  algorithmic bytes per launch: {args.alg_bytes / 1e9:.3f} GB -> traffic / algorithmic = "
                  f"{tr / args.alg_bytes:.3f}; algorithmic GB/s under ncu = {args.alg_bytes / t_ns:.1f}")
    for name, H, body in source_tables(args.rep):
        if args.kernel not in name:
            continue
        si, ei, sa = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
        stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
        stalls, ops, samp = collections.Counter(), collections.Counter(), collections.Counter()
        tot = stot = 0.0
        for r in body:
            try:
                n, s = float(r[ei]), float(r[sa])
            except (ValueError, IndexError):
                continue
            for i in stall_cols:
                try:
                    stalls[H[i]] += float(r[i])
                except ValueError:
                    pass
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
            if m:
                ops[m.group(2)] += n
                samp[m.group(2)] += s
                tot += n
                stot += s
        ssum = sum(stalls.values()) or 1.0
        print("\n  warp-state samples (all): " +
              "  ".join(f"{k[6:]}={100 * v / ssum:.1f}%" for k, v in stalls.most_common(9)))
        per = f", {tot / args.rows:.1f} per row" if args.rows else ""
        print(f"\n  SASS opcode mix ({tot:.3e} warp instructions{per}):")
        for op, n in ops.most_common(28):
            extra = f" {n / args.rows:8.1f}/row" if args.rows else ""
            print(f"    {op:12s} {100 * n / tot:6.2f}%{extra}   samples {100 * samp[op] / max(stot, 1):5.1f}%")
        break


if __name__ == "__main__":
    main()
