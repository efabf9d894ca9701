"""Writes tests/golden/matlab_ply_write.ply: a small cloud in the byte layout matlab/utils/ply/ply_write.m emits in
'ascii' mode for the struct of write_ply_cloud.m:37-53 (header text ply_write.m:95,118,197,204; values '%-.6f ',
ply_write.m:89,225) -- written with plain string formatting, independent of objective_slam_b200.io -- plus the
`.trans_adj` side file of compute_normals.m:17-22 and the values as .npy for the reader test."""
import os
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(0xD205)
pts = rng.uniform(-40, 60, size=(12, 3))
nrm = rng.normal(size=(12, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
adj = np.abs(pts.min(0)) + 1.0                                     # compute_trans_adj.m:8-10
pts = pts + adj                                                    # compute_normals.m:11-13
with open(os.path.join(here, "matlab_ply_write.ply"), "w", newline="\n") as f:
    f.write("ply\nformat ascii 1.0\ncomment created by MATLAB ply_write\nelement vertex %u\n" % len(pts))
    for n in ("x", "y", "z", "nx", "ny", "nz"):
        f.write("property float %s\n" % n)
    f.write("end_header\n")
    for a, b in zip(pts, nrm):
        for v in (*a, *b):
            f.write("%-.6f " % v)
        f.write("\n")
with open(os.path.join(here, "matlab_ply_write.ply.trans_adj"), "w", newline="\n") as f:
    f.write("%f %f %f\n" % tuple(adj))
np.save(os.path.join(here, "matlab_ply_write_values.npy"), np.concatenate([pts, nrm], 1))
