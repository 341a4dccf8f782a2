# per-kernel durations of a few config-5 iterations (8 x 256^3, stepped engine) under ncu
mkdir -p gpurun_out/ncu_r2
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pcg_phase|sptrsv_ts|arm_positions" -c 400 --csv --log-file gpurun_out/ncu_r2/launches_c5.csv python bench.py --config c5 --c5-per-gpu 8 --steps 1 --warmup 1 --max-iter 12 > gpurun_out/ncu_r2/c5_under_ncu.log 2>&1
tail -2 gpurun_out/ncu_r2/c5_under_ncu.log | cut -c1-300
python - <<PY
import csv
rows=list(csv.reader(l for l in open("gpurun_out/ncu_r2/launches_c5.csv") if l.startswith('"')))
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
tot={}
for r in rows[1:]:
    k=r[ki][:70]; t=float(r[vi].replace(",","")); a=tot.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=t
for k,(n,t) in sorted(tot.items(), key=lambda kv:-kv[1][1]): print(f"{n:4d} x {t/n/1e3:9.1f} us  {k}")
PY
