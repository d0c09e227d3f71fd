"""Development aid: per-role digest of an ncu report of mlp_chain_kernel (source page): samples, issue share and top stall reasons for the
SASS ranges of the producer / MMA issuer / epilogue code, and the hottest instructions.  usage: ncu_roles.py report.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"):
    for h, v in zip(hdr, vals):
        if h == k: print(k, v)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
utc = [i for i, r in enumerate(data) if "UTCHMMA" in r[ia]]
tma = [i for i, r in enumerate(data) if "UTMALDG" in r[ia]]
ldtm = [i for i, r in enumerate(data) if "LDTM" in r[ia]]
def agg(name, a, b):
    tot = {}
    for r in data[a:b]:
        for i in cols: tot[hdr[i]] = tot.get(hdr[i], 0) + int(r[i] or 0)
    s = sum(tot.values())
    print(name, a, b, "samples", s, sorted(((v, k) for k, v in tot.items() if v), reverse=True)[:6])
print("total samples", sum(int(r[isamp]) for r in data))
if utc: agg("mma-issuers", min(utc) - 120, max(utc) + 60)
if tma: agg("producers", min(tma) - 60, max(tma) + 60)
if ldtm: agg("epilogue", min(ldtm) - 60, max(ldtm) + 400)
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:25]
for i in sorted(top): print(i, data[i][isamp], data[i][iex], data[i][ia][:90])
