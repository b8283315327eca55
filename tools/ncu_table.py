"""One line per kernel launch from an `ncu --set full --page raw --csv` dump: duration, DRAM traffic and utilisation, tensor pipe,
issue slots, shared-memory wavefronts (LSU side; tensor-core operand reads are not counted there), achieved occupancy."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"s": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "second": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3,  # -> microseconds
         "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}                  # -> megabytes


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", "")) * SCALE.get(units[ix[k]], 1.0)
    except (KeyError, ValueError):
        return float("nan")


print(f"{'kernel':46s} {'us':>8s} {'dramR_MB':>8s} {'dramW_MB':>8s} {'dram%':>6s} {'tensor%':>7s} {'issue%':>6s} {'l1tex%':>6s} {'smemWf_M':>8s} {'warps%':>6s}")
for r in rows[2:]:
    name = r[ix["Kernel Name"]].replace("epnet::", "")[:46]
    print(f"{name:46s} {f(r, 'gpu__time_duration.sum'):8.1f} {f(r, 'dram__bytes_read.sum'):8.1f} {f(r, 'dram__bytes_write.sum'):8.1f} "
          f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{f(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{f(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6:8.2f} "
          f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f}")
