for a in "32x32x32 64 28" "32x32x32 64 8" "32x32x32 64 28 sparse" "16x16x16x16 16 8" "64x64 4096 28 sparse"; do ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bwd_raw.csv python scratch/bwd_ops.py $a > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/bwd_raw.csv")) if len(r)>14 and r[0].isdigit()]
print("$a", " ".join(f"{float(r[14])/1e3:.1f}" for r in rows[-19:] if "convnd_tc_kernel<2" in r[4]))
PY
done
