"""STL part -> axis-aligned box obstacles (an extension of the reference).

The reference only loads its STL maps for display: Lib/functions/MapFromSTL.m:1-12 reads map/assembly line_Assem1.stl with
stlread, shifts every coordinate so that its minimum is 0, lowers the second one by 100 (mm), and permutes the axes
(x, y, z) <- (z, x, y); the result (`envir`, in mm) is what map/environment.mat stores and DrawMap patches; the distance
function that would have used it (point2surface_dis, Lib/functions/dist_arm_surface.m:44) is defined nowhere.  Here the same
transformed triangles (converted to metres, the robot's unit) are covered by a small set of axis-aligned boxes -- a k-d split of
the triangles at the median centroid along the longest axis -- which libcfs_b200 takes as obstacles of kind 'box'
(cfs_set_obstacles_ex): obs{j}.shape = 'box', obs{j}.l = [min corner, max corner].
"""
import numpy as np


def read_stl(path):
    """binary STL -> triangles (n, 3, 3) float64, file units"""
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw[80:84], dtype="<u4")[0])
    if len(raw) < 84 + 50 * n:
        raise ValueError("%s: not a binary STL (header says %d triangles, file has %d bytes)" % (path, n, len(raw)))
    rec = np.frombuffer(raw[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    return rec["v"].astype(np.float64)


def map_from_stl(tri_mm):
    """the vertex transform of Lib/functions/MapFromSTL.m:6-10, then mm -> m"""
    v = tri_mm.reshape(-1, 3).copy()
    v -= v.min(axis=0)                      # :6-8
    v[:, 1] -= 100.0                        # :9
    v = v[:, [2, 0, 1]]                     # :10-11  (x, y, z) <- (z, x, y)
    return (v / 1000.0).reshape(-1, 3, 3)


def boxes_from_triangles(tri, max_boxes=16):
    """k-d cover: split the triangle set at the median centroid along the longest axis of its bounding box until max_boxes
    leaves; every leaf's box is the bounding box of ITS triangles (so the union of the boxes contains the whole mesh)."""
    leaves = [tri]
    while len(leaves) < max_boxes:
        k = int(np.argmax([np.prod(np.maximum(t.reshape(-1, 3).max(0) - t.reshape(-1, 3).min(0), 1e-9)) if len(t) > 1 else -1 for t in leaves]))
        t = leaves[k]
        if len(t) < 2:
            break
        v = t.reshape(-1, 3)
        ax = int(np.argmax(v.max(0) - v.min(0)))
        c = t.mean(axis=1)[:, ax]
        med = np.median(c)
        left, right = t[c <= med], t[c > med]
        if len(left) == 0 or len(right) == 0:
            break
        leaves[k:k + 1] = [left, right]
    return [(t.reshape(-1, 3).min(0), t.reshape(-1, 3).max(0)) for t in leaves]


def boxes_from_stl(path, max_boxes=16, D=0.05, epsilon=0.05, near=None, radius=None, max_keep=None, max_size=None):
    """obs list for Context.set_obstacles / CFS_FANUC: the boxes of the STL part (MapFromSTL transform, metres).  near / radius:
    keep only the boxes whose closest point is within radius of `near` (e.g. the robot base); max_keep: the closest ones;
    max_size: drop leaves whose largest extent exceeds it (floor slabs and walls of a whole-cell mesh, which would contain the
    robot itself)."""
    boxes = boxes_from_triangles(map_from_stl(read_stl(path)), max_boxes)
    if max_size is not None:
        boxes = [(lo, hi) for lo, hi in boxes if (hi - lo).max() <= max_size]
    if near is not None:
        near = np.asarray(near, dtype=np.float64)
        dist = [float(np.linalg.norm(np.maximum(np.maximum(lo - near, near - hi), 0.0))) for lo, hi in boxes]
        order = np.argsort(dist, kind="stable")
        boxes = [boxes[i] for i in order if radius is None or dist[i] <= radius]
        if max_keep is not None:
            boxes = boxes[:max_keep]
    return [{"shape": "box", "l": np.stack([lo, hi], axis=1), "D": D, "epsilon": epsilon} for lo, hi in boxes]
