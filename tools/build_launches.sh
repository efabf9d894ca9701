# per-kernel times of one 10k-point model table build (second build of the process), own sort and library sort
cat > /tmp/b.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import objective_slam_b200 as ppf
from objective_slam_b200 import synth
mp, mn = synth.make_model(int(sys.argv[1]), seed=0xD205 + 3)
d = synth.d_dist_for(mp)
for i in range(2):
    m = ppf.Model(mp, mn, d); m.close()
PY
for mode in own cub; do
  PPF_B200_SORT=$mode ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/build_${mode}.csv python /tmp/b.py 10000 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/build_${mode}.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value")
data=[(r[kn].split("(")[0][:60], float(r[mv].replace(",",""))/1e3) for r in rows[hi+1:] if len(r)>mv]
half=len(data)//2
print("== ${mode}: second build, us per launch")
tot=0
for n,t in data[half:]:
    print(f"  {t:9.1f}  {n}"); tot+=t
print(f"  total {tot:.1f} us")
PY
done
