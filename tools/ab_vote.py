"""A/B timing of vote-kernel variants: each library under objective_slam_b200/lib_ab/<name>/ runs the bench workload
(configs[1]: 10k-point model, 50k-point scene, ref_point_df 8) in its own process (PPF_B200_LIB selects the build).
usage: ab_vote.py name [name ...]   (optional env AB_ARGS="n_model n_scene df reps")"""
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = os.environ.get("AB_ARGS", "10000 50000 8 4").split()
for name in sys.argv[1:]:
    lib = os.path.join(root, "objective_slam_b200", "lib_ab", name, "libppf_b200.so")
    env = dict(os.environ, PPF_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_vote.py"), *args], env=env,
                       capture_output=True, text=True)
    ms = [float(l.split("ms_vote")[1].split()[0]) for l in r.stdout.splitlines() if "ms_vote" in l]
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    print(f"{name:12s} best ms_vote {min(ms) if ms else float('nan'):9.3f}  all {ms}  | {tail}", flush=True)
