"""Developer tool: condense an .ncu-rep into the text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu.txt
"""

import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of {rep}  (ncu --set full --clock-control none --import-source on)")
    for r in data:
        print(f"\n## kernel: {r[ci['Kernel Name']]}  id={r[ci['ID']]}")
        for k in KEYS:
            if k in ci:
                print(f"{k:75s} {r[ci[k]]:>16s} {units[ci[k]]}")
        print("warp stall reasons (per issue-active):")
        for h in hdr:
            if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h:
                v = float(r[ci[h]] or 0)
                if v >= 0.05:
                    print(f"  {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")
    sass = page(rep, "source", ["--print-source", "sass"])
    if len(sass) > 2:
        h = sass[1]
        c = {x: i for i, x in enumerate(h)}
        body = sass[2:]
        tot = sum(int(r[c["# Samples"]]) for r in body)
        print(f"\n## hottest SASS instructions (of {tot} samples)")
        for r in sorted(body, key=lambda r: -int(r[c["# Samples"]]))[:16]:
            st = {k: int(r[c[k]]) for k in h if k.startswith("stall_") and "Not Issued" not in k and int(r[c[k]]) > 0}
            top = max(st.items(), key=lambda kv: kv[1])[0] if st else ""
            print(f"  {int(r[c['# Samples']]):7d}  {r[c['Source']].strip()[:64]:64s} {top}")
        agg = {}
        for r in body:
            op = (r[c["Source"]].strip().split() or [""])[0]
            if op.startswith("@"):
                op = (r[c["Source"]].strip().split() + [""])[1]
            w, wi = int(r[c["L1 Wavefronts Shared"]]), int(r[c["L1 Wavefronts Shared Ideal"]])
            if w:
                a = agg.setdefault(op, [0, 0])
                a[0] += w
                a[1] += wi
        print("\n## shared-memory wavefronts by opcode (actual / ideal)")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:12s} {v[0]:12d} / {v[1]:12d}")


if __name__ == "__main__":
    main()
