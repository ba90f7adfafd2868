"""Summarise an .ncu-rep (read on the CPU box): per kernel duration, DRAM bytes / throughput, tensor-pipe activity,
achieved fraction of the measured peaks.  usage: ncu_summary.py file.ncu-rep [out.txt]"""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]
def f(r, k):
    if k not in col: return None
    try: return float(r[col[k]].replace(",", ""))
    except ValueError: return None
def scale(k, v):
    if v is None: return None
    u = units[col[k]]
    m = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    return v * m.get(u, 1.0)
out = []
for r in data:
    name = r[col["Kernel Name"]][:70]
    dur = scale("gpu__time_duration.sum", f(r, "gpu__time_duration.sum"))
    rd, wr = scale("dram__bytes_read.sum", f(r, "dram__bytes_read.sum")), scale("dram__bytes_write.sum", f(r, "dram__bytes_write.sum"))
    tens = f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    gbs = (rd + wr) / dur / 1e9 if dur and rd is not None else None
    out.append({"kernel": name, "grid": r[col["launch__grid_size"]] if "launch__grid_size" in col else None, "us": dur * 1e6 if dur else None,
                "dram_read_MB": rd / 1e6 if rd is not None else None, "dram_write_MB": wr / 1e6 if wr is not None else None,
                "dram_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peaks["hbm_gbs"] if gbs else None,
                "dram_pct_ncu": f(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"), "tensor_pipe_active_pct": tens,
                "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), "regs": f(r, "launch__registers_per_thread")})
txt = "\n".join(json.dumps(o) for o in out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
