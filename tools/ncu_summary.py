"""Text summary of an ncu report (one block per captured launch) for profiles/:
    python tools/ncu_summary.py gpurun_out/r02_sort.ncu-rep [--hbm-peak 6546.6] [--opcodes] > profiles/r02_sort_ncu_summary.txt
Per launch: duration, DRAM bytes read / written, achieved DRAM GB/s and its fraction of the measured HBM peak
(MEASURED_PEAKS.json), occupancy, registers, shared memory, issue utilisation, pipe utilisation.
--opcodes adds the SASS opcode mix (warp instructions) of the FIRST launch from the source page."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram written"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving / issue"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
        "s": 1.0, "second": 1.0, "nsecond": 1e-9}


def main():
    rep = sys.argv[1]
    peak = 6546.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if "--hbm-peak" in sys.argv:
        peak = float(sys.argv[sys.argv.index("--hbm-peak") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    print(f"# {os.path.basename(rep)}: {len(rows) - 2} captured launch(es); HBM peak {peak} GB/s (MEASURED_PEAKS.json)")
    body = rows[2:]
    if "--by-name" in sys.argv:
        # one line per kernel (template arguments shortened): launches, total time, DRAM bytes, GB/s; then the full block of the
        # longest launch of each of the top kernels only
        def num(r, key):
            if key not in ix or r[ix[key]] == "":
                return 0.0
            return float(r[ix[key]].replace(",", "")) * UNIT.get(units[ix[key]], 1.0)
        import re
        def short(name):
            m = re.search(r"agg::(Tag\w+)", name)
            if m:
                return ("agg_kernel_if<" if "agg_kernel_if" in name else "agg_kernel<") + m.group(1) + ">"
            return name.split("(")[0][:60]
        agg = collections.OrderedDict()
        for r in body:
            k = short(r[ix["Kernel Name"]])
            a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, None, 0.0])
            t = num(r, "gpu__time_duration.sum")
            a[0] += 1
            a[1] += t
            a[2] += num(r, "dram__bytes_read.sum")
            a[3] += num(r, "dram__bytes_write.sum")
            if t > a[5]:
                a[4], a[5] = r, t
        tot = sum(a[1] for a in agg.values())
        print(f"# by kernel (total {tot * 1e3:.3f} ms under ncu: cold caches, serialised)")
        print(f"# {'kernel':44s} {'launches':>8s} {'ms':>9s} {'share':>6s} {'DRAM MB':>9s} {'GB/s':>8s} {'of peak':>7s}")
        top = sorted(agg.items(), key=lambda kv: -kv[1][1])
        for k, a in top:
            gbs = (a[2] + a[3]) / a[1] / 1e9 if a[1] > 0 else 0.0
            print(f"  {k:44s} {a[0]:8d} {a[1] * 1e3:9.3f} {a[1] / tot:6.3f} {(a[2] + a[3]) / 1e6:9.1f} {gbs:8.1f} {gbs / peak:7.3f}")
        body = [a[4] for k, a in top[:6]]
        print()
        print("# full metrics of the longest launch of each of the six kernels with the most time:")
    for r in body:
        print()
        print(r[ix["Kernel Name"]][:150])

        def val(key):
            if key not in ix or r[ix[key]] == "":
                return None, ""
            return float(r[ix[key]].replace(",", "")), units[ix[key]]
        dur, du = val("gpu__time_duration.sum")
        rd, ru = val("dram__bytes_read.sum")
        wr, wu = val("dram__bytes_write.sum")
        if dur is not None and rd is not None and wr is not None:
            sec = dur * UNIT.get(du, 1.0)
            gbs = (rd * UNIT.get(ru, 1.0) + wr * UNIT.get(wu, 1.0)) / sec / 1e9
            print(f"    achieved DRAM throughput           {gbs:10.1f} GB/s = {gbs / peak:.3f} of the measured HBM peak")
        for key, label in KEYS:
            v, u = val(key)
            if v is not None:
                print(f"    {label:38s} {v:14.3f} {u}")
    if "--opcodes" in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        h = srows[1]
        ie, so = h.index("Instructions Executed"), h.index("Source")
        ops = collections.Counter()
        tot = 0
        for r in srows[2:]:
            try:
                n = int(r[ie])
            except Exception:
                continue
            t = r[so].split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ops[op.split(".")[0]] += n
            tot += n
        print()
        print(f"SASS opcode mix of the first launch (warp instructions, total {tot}):")
        print("    " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(24)))


if __name__ == "__main__":
    main()
