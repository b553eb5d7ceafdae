#!/usr/bin/env bash
mkdir -p gpurun_out
echo "== launch list of the GEMM experiment (kernel-only durations)"
TG_GEMM_STREAMK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_x3" --csv --log-file gpurun_out/j5_launches.csv python tools/exp_gemm2.py > gpurun_out/j5_l.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/j5_launches.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; ki, gi, vi = H.index("Kernel Name"), H.index("Grid Size"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    key = (("3-product" if "Lb1ELb1" in r[ki] or "true, true" in r[ki] else "4-mult") , r[gi])
    agg.setdefault(key, []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    v = sorted(v)
    print(k, "n=%d median %.1f us min %.1f us" % (len(v), v[len(v)//2] / 1e3, v[0] / 1e3))
PY
echo "== full capture: 3-product stream-K on C2 shape, and the M=128 shard (4-mult)"
TG_GEMM_STREAMK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_x3" -s 6 -c 2 -o gpurun_out/j5_gemm3_c2 python tools/exp_gemm2.py > gpurun_out/j5_p1.log 2>&1
ls -la gpurun_out/*.ncu-rep
