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
    for r in rows[2:]:
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
