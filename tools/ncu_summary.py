"""Summarise an .ncu-rep: headline raw metrics per kernel + hottest SASS lines by stall samples.
usage: python tools/ncu_summary.py report.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("==", name[:100])
    for h, u, v in zip(hdr, units, r):
        if h in KEYS:
            print(f"   {h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name",')
for b in blocks[1:]:
    lines = list(csv.reader(io.StringIO(b)))
    kname = lines[0][0]
    hdr = lines[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in lines[2:] if len(r) == len(hdr)]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, "# Samples") for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((k, sum(f(r, k) for r in data)) for k in stalls), key=lambda kv: -kv[1])[:7]
    print(f"== source: {kname[:80]}  samples={tot:.0f} inst={sum(f(r, 'Instructions Executed') for r in data):.0f}")
    print("   stalls:", ", ".join(f"{k[6:]}={v / max(tot, 1) * 100:.0f}%" for k, v in agg))
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
        st = sorted(((k, f(r, k)) for k in stalls), key=lambda kv: -kv[1])[:2]
        print(f"   {f(r, '# Samples') / max(tot, 1) * 100:5.1f}%  x{f(r, 'Instructions Executed'):11.0f}  "
              f"{r[ix['Source']][:60]:60s} {st[0][0][6:]}:{st[0][1]:.0f} {st[1][0][6:]}:{st[1][1]:.0f}")
