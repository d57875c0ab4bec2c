"""Turn an ncu report (--set full) into the table the roofline discussion needs: per kernel launch the
duration, DRAM bytes and % of peak DRAM throughput, tensor-pipe active %, L1/L2 throughput %, achieved
occupancy, registers -- and the achieved GB/s against MEASURED_PEAKS.json.
    python bench/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md      (runs on the CPU box)"""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))


def get(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit, 1)


print("| kernel | grid x block | time us | DRAM read+write MB | achieved GB/s (%% of the measured %.0f GB/s) | dram %% | tensor pipe %% | l1tex %% | L2 %% | achieved occ %% | regs |"
      % peaks.get("hbm_gbs", 6567.0))
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in data:
    name = get(r, "Kernel Name")
    t = to_us(get(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]])
    rd = to_bytes(get(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
    wr = to_bytes(get(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
    gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else float("nan")
    print("| `%s` | %s x %s | %.1f | %.1f | %.0f (%.0f%%) | %s | %s | %s | %s | %s | %s |" % (
        name[:90], get(r, "launch__grid_size"), get(r, "launch__block_size"), t, (rd + wr) / 1e6, gbs,
        100 * gbs / peaks.get("hbm_gbs", 6567.0), get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", get(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")),
        get(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"), get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), get(r, "launch__registers_per_thread")))
