"""Summarise an `ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv` launch list: the LAST denoiser evaluation in the
file (from its embed_kernel launch on), per launch and in total.  Usage: python scripts/ncu_metrics_summary.py file.csv"""
import collections
import csv
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "%": 1}
TENSOR = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"


def load(path):
    rows = [r for r in csv.reader(ln for ln in open(path) if ln.startswith('"'))]
    head, out = rows[0], collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(head, r))
        e = out.setdefault(int(d["ID"]), {"name": d["Kernel Name"], "grid": d["Grid Size"]})
        e[d["Metric Name"]] = float(d["Metric Value"].replace(",", "")) * SCALE.get(d["Metric Unit"], 1)
    return list(out.values())


def main(path):
    ks = load(path)
    start = [i for i, k in enumerate(ks) if "embed_kernel" in k["name"]][-1]
    ev = [k for k in ks[start:] if k["name"].startswith(("b2p::", "void b2p::"))]
    print(f"# {path}: last evaluation = {len(ev)} launches")
    print(f"# {'kernel':<34}{'grid':<14}{'us':>8}{'L2 MB':>10}{'L2 GB/s':>10}{'DRAM MB':>10}{'tensor %':>10}")
    for k in ev:
        t, l2 = k["gpu__time_duration.sum"], k["lts__t_bytes.sum"]
        dr = k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]
        print(f"  {k['name'].replace('void ', '')[:32]:<34}{k['grid']:<14}{t / 1e3:8.2f}{l2 / 1e6:10.2f}{l2 / t:10.0f}{dr / 1e6:10.3f}{k.get(TENSOR, 0.0):10.1f}")
    T = sum(k["gpu__time_duration.sum"] for k in ev)
    L = sum(k["lts__t_bytes.sum"] for k in ev)
    D = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in ev)
    tc = [k for k in ev if "conv_tc_kernel" in k["name"]]
    print(f"# total: {T / 1e3:.1f} us of kernel time (serialised under ncu), L2 {L / 1e6:.1f} MB ({L / T:.0f} GB/s), DRAM {D / 1e6:.2f} MB ({D / T:.1f} GB/s)")
    if tc:
        tt = sum(k["gpu__time_duration.sum"] for k in tc)
        print(f"# conv_tc_kernel: {len(tc)} launches, {tt / 1e3:.1f} us, time-weighted tensor-pipe active "
              f"{sum(k.get(TENSOR, 0.0) * k['gpu__time_duration.sum'] for k in tc) / tt:.1f} %, "
              f"L2 {sum(k['lts__t_bytes.sum'] for k in tc) / tt:.0f} GB/s")


if __name__ == "__main__":
    main(sys.argv[1])
