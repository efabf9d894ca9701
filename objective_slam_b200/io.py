"""Host-side I/O around the hot path: PLY clouds, ground-truth poses, and the
reference's validation metric.  Plain numpy; nothing here is on the hot path.

* PLY subset: `ply` / `format ascii|binary_little_endian|binary_big_endian 1.0`, one `vertex` element with
  float/double (or integer) scalar properties, among them `x y z` and either `nx ny nz` (what
  matlab/write_ply_cloud.m:37-53 writes through ply_write.m) or `normal_x normal_y normal_z` (PCL's names,
  alignment.cpp:212,241 reads them with pcl::io::loadPLYFile<PointNormal>).  Other elements (faces) are skipped.
* `.trans_adj` side files (compute_trans_adj.m:2-16, compute_normals.m:11-22): the per-dataset shift into the positive
  octant that the reference's pipeline applies before recognition, and the pose bookkeeping that goes with it.
* Poses: 4x4 row-major text matrices, the format alignment.cpp:304-307 reads through util.hpp:95-104.
* ht_dist / validation: linalg.cu:9-20 and alignment.cpp:317-323.
"""
from __future__ import annotations

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
    "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
    "double": "f8", "float64": "f8",
}


def read_ply(path):
    """Returns (points[N,3] float32, normals[N,3] float32 or None)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    cur["props"].append(("list", tok[2], tok[3], tok[4]))
                else:
                    cur["props"].append((tok[1], tok[2]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        vertex = None
        for el in elements:
            if el["name"] != "vertex":
                if vertex is not None:
                    break                                     # everything we need has been read
                if any(p[0] == "list" for p in el["props"]):
                    raise ValueError(f"{path}: a list element precedes the vertex element")
                # skip a fixed-size element that precedes the vertices
                if fmt == "ascii":
                    for _ in range(el["count"]):
                        f.readline()
                else:
                    f.seek(sum(np.dtype(_PLY_TYPES[t]).itemsize for t, _ in el["props"]) * el["count"], 1)
                continue
            if any(p[0] == "list" for p in el["props"]):
                raise ValueError(f"{path}: list property in the vertex element")
            names = [n for _, n in el["props"]]
            if fmt == "ascii":
                rows = [f.readline().split() for _ in range(el["count"])]
                data = np.array(rows, dtype=np.float64).reshape(el["count"], len(names))
                cols = {n: data[:, i] for i, n in enumerate(names)}
            else:
                end = "<" if fmt == "binary_little_endian" else ">"
                dt = np.dtype([(n, end + _PLY_TYPES[t]) for t, n in el["props"]])
                rec = np.frombuffer(f.read(dt.itemsize * el["count"]), dtype=dt, count=el["count"])
                cols = {n: rec[n] for n in names}
            vertex = cols
        if vertex is None:
            raise ValueError(f"{path}: no vertex element")
    pts = np.stack([vertex["x"], vertex["y"], vertex["z"]], 1).astype(np.float32)
    for trio in (("nx", "ny", "nz"), ("normal_x", "normal_y", "normal_z")):
        if all(k in vertex for k in trio):
            return pts, np.stack([vertex[k] for k in trio], 1).astype(np.float32)
    return pts, None


def write_ply(path, points, normals=None, fmt="ascii", pcl_names=False):
    """Writes what write_ply_cloud.m:37-53 writes: vertex x y z nx ny nz (float), ascii by default."""
    p = np.asarray(points, np.float32)
    cols = [p]
    names = ["x", "y", "z"]
    if normals is not None:
        cols.append(np.asarray(normals, np.float32))
        names += ["normal_x", "normal_y", "normal_z"] if pcl_names else ["nx", "ny", "nz"]
    data = np.concatenate(cols, 1).astype(np.float32)
    header = ["ply", f"format {fmt} 1.0", "comment written by objective_slam_b200", f"element vertex {len(p)}"]
    header += [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        if fmt == "ascii":
            for row in data:
                f.write((" ".join(repr(float(np.float32(v))) for v in row) + "\n").encode("ascii"))
        elif fmt == "binary_little_endian":
            f.write(data.astype("<f4").tobytes())
        elif fmt == "binary_big_endian":
            f.write(data.astype(">f4").tobytes())
        else:
            raise ValueError(fmt)


def write_ply_matlab(path, points, normals):
    """Byte layout of matlab/utils/ply/ply_write.m in 'ascii' mode for the struct write_ply_cloud.m:37-53 and
    compute_normals.m:6-15 build: its comment line, `property float` x y z nx ny nz (doubles are narrowed to float
    unless 'double' is asked for, ply_write.m:192-194), every value printed with '%-.6f ' (ply_write.m:89,225), i.e.
    six decimals and a trailing blank before the newline."""
    p = np.asarray(points, np.float64)
    n = np.asarray(normals, np.float64)
    with open(path, "w", newline="\n") as f:
        f.write("ply\nformat ascii 1.0\ncomment created by MATLAB ply_write\n")
        f.write(f"element vertex {len(p)}\n")
        for name in ("x", "y", "z", "nx", "ny", "nz"):
            f.write(f"property float {name}\n")
        f.write("end_header\n")
        for a, b in zip(p, n):
            f.write("".join("%-.6f " % v for v in (*a, *b)) + "\n")


# ---- .trans_adj: the shift that moves every cloud of a dataset into the positive octant ---------------------
def compute_trans_adj(clouds):
    """compute_trans_adj.m:2-16 over a list of point arrays: per axis max over the clouds of |min| + 1."""
    t = np.zeros(3)
    for pts in clouds:
        t = np.maximum(t, np.abs(np.asarray(pts, np.float64).min(axis=0)) + 1.0)
    return t


def write_trans_adj(ply_path, trans_adj):
    """compute_normals.m:17-22: `<output>.trans_adj`, one line '%f %f %f'."""
    with open(str(ply_path) + ".trans_adj", "w", newline="\n") as f:
        f.write("%f %f %f\n" % tuple(float(v) for v in np.asarray(trans_adj).reshape(3)))


def read_trans_adj(ply_path):
    """The shift stored next to a cloud by compute_normals.m, or None when the cloud has none."""
    import os
    path = str(ply_path) + ".trans_adj"
    if not os.path.exists(path):
        return None
    v = np.loadtxt(path, dtype=np.float64).reshape(-1)
    if v.size != 3:
        raise ValueError(f"{path}: expected 3 numbers, got {v.size}")
    return v


def apply_trans_adj(points, trans_adj):
    """compute_normals.m:11-13: points + trans_adj (the reference's pipeline wants positive coordinates)."""
    return (np.asarray(points, np.float64) + np.asarray(trans_adj, np.float64).reshape(1, 3)).astype(np.float32)


def pose_in_adjusted_frame(T, trans_adj_model, trans_adj_scene):
    """A model->scene pose expressed between the ORIGINAL clouds, rewritten for clouds shifted by their .trans_adj:
    x' = x + a_m, y' = y + a_s  =>  T' = Trans(a_s) T Trans(-a_m)."""
    A = np.eye(4); A[:3, 3] = np.asarray(trans_adj_scene, np.float64).reshape(3)
    B = np.eye(4); B[:3, 3] = -np.asarray(trans_adj_model, np.float64).reshape(3)
    return A @ np.asarray(T, np.float64).reshape(4, 4) @ B


def read_pose(path):
    """4x4 ground-truth matrix (alignment.cpp:304-307): 16 whitespace-separated numbers, row-major."""
    v = np.loadtxt(path, dtype=np.float64).reshape(-1)
    if v.size != 16:
        raise ValueError(f"{path}: expected 16 numbers, got {v.size}")
    return v.reshape(4, 4)


def write_pose(path, T):
    np.savetxt(path, np.asarray(T, np.float64).reshape(4, 4), fmt="%.9g")


def ht_dist(a, b):
    """linalg.cu:9-20: (norm of the translation difference, |angle| of a_rot^-1 * b_rot)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    dt = float(np.linalg.norm(a[:3, 3] - b[:3, 3]))
    R = a[:3, :3].T @ b[:3, :3]
    ang = float(abs(np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0))))
    return dt, ang


def validate_pose(estimate, truth, model_diameter, translation_threshold=0.1, rotation_threshold_deg=12.0):
    """alignment.cpp:317-323: 1 if |dt| < translation_threshold * diameter and angle < rotation_threshold."""
    dt, ang = ht_dist(estimate, truth)
    return int(dt < translation_threshold * model_diameter and ang < np.radians(rotation_threshold_deg)), dt, ang


def model_diameter(points):
    """Largest extent of the axis-aligned bounding box, the normaliser of alignment.cpp:246-253."""
    p = np.asarray(points, np.float64)
    return float((p.max(0) - p.min(0)).max())
