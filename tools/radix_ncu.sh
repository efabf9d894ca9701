cat > /tmp/b.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import objective_slam_b200 as ppf
from objective_slam_b200 import synth
mp, mn = synth.make_model(10000, seed=0xD205 + 3)
d = synth.d_dist_for(mp)
for i in range(2):
    m = ppf.Model(mp, mn, d); m.close()
PY
ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 2 -c 2 -o gpurun_out/radix_scatter -f python /tmp/b.py > gpurun_out/radix_ncu.log 2>&1
tail -3 gpurun_out/radix_ncu.log
