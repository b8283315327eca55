"""One line per kernel launch from an `ncu --page raw --csv` dump: time, DRAM/L2 traffic and throughput, tensor pipe, occupancy."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def g(r, k, default=""):
    return r[ix[k]] if k in ix else default


keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
print("available tensor metrics:", [h for h in hdr if "tensor" in h][:12])
for r in rows[2:]:
    name = g(r, "Kernel Name")[:46]
    vals = []
    for k in keys:
        if k in ix:
            vals.append("%s=%s%s" % (k.split(".")[0].replace("__", ":"), r[ix[k]], units[ix[k]] and " " + units[ix[k]]))
    print(name)
    print("    " + "; ".join(vals))
