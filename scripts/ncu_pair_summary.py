"""Summary table of an `ncu --set full -k regex:gemm_pair_kernel` capture of one eager step (raw CSV page).
Usage: python scripts/ncu_pair_summary.py <raw.csv> [title]"""
import csv
import sys

NAMES = ["QE input projection", "prop_fc fwd (+gate)", "conv0 fwd (K-split 2)", "FPN inner x3 fwd", "FPN layer x3 fwd", "towers x3 fwd",
         "mix_fc x3 fwd", "iou_scores.0 x3 fwd", "towers bwd (dgrad+wgrad)", "FPN layer bwd", "FPN inner bwd", "conv2 bwd", "conv1 bwd",
         "conv0 bwd", "prop_fc wgrad", "QE bwd projections"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
c = lambda n: hdr.index(n)  # noqa: E731
tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
print(sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none, gemm_pair_kernel launches of one (eager) step")
print("%-28s %6s %9s %9s %10s %10s %9s %9s %8s" % ("launch", "CTAs", "us", "SM MHz", "DRAM rd MB", "DRAM wr MB", "L2->SM GB", "tensor %", "lts %"))
for i, r in enumerate(data[:16]):
    grid = int(r[c("Grid Size")].strip("()").split(",")[0])
    us = float(r[c("gpu__time_duration.sum")])
    mhz = float(r[c("sm__cycles_elapsed.max")]) / us
    print("%-28s %6d %9.1f %9.0f %10.1f %10.1f %9.2f %9.1f %8.1f" % (
        NAMES[i], grid, us, mhz, float(r[c("dram__bytes_read.sum")]), float(r[c("dram__bytes_write.sum")]),
        float(r[c("lts__t_sectors_srcunit_tex_op_read.sum")]) * 32 / 1e9, float(r[c(tp)]),
        float(r[c("lts__throughput.avg.pct_of_peak_sustained_elapsed")])))
