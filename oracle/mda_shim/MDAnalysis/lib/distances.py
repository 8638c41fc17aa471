"""TEST INFRASTRUCTURE ONLY -- restated `MDAnalysis.lib.distances.capped_distance`.

Brute force only: the reference always passes a single reference coordinate
(`/root/reference/waterEntropy/maths/trig.py:26-34`), for which MDAnalysis'
`_determine_method` picks `bruteforce` (len(reference) < 10).

Orthorhombic arithmetic follows `calc_distances.h::_calc_distance_array_ortho`
+ `minimum_image`: the coordinate difference is taken
in float32 and widened, `s = (double)inv_box_f32 * dx`,
`dx = (double)box_f32 * (s - round(s))`, `d = sqrt(dx0^2 + dx1^2 + dx2^2)`.
Triclinic boxes: coordinates wrapped into the cell, then a 27-image search
over the lower-triangular box vectors (parity UNPINNED by any reference test).
"""
import numpy as np


def triclinic_vectors(dimensions):
    """Lower-triangular float32 box matrix from [lx,ly,lz,alpha,beta,gamma]."""
    dim = np.asarray(dimensions, dtype=np.float32)
    lx, ly, lz, alpha, beta, gamma = [np.float64(v) for v in dim]
    box = np.zeros((3, 3), dtype=np.float32)
    if alpha == 90.0 and beta == 90.0 and gamma == 90.0:
        box[0, 0], box[1, 1], box[2, 2] = lx, ly, lz
        return box
    cos_a = 0.0 if alpha == 90.0 else np.cos(np.deg2rad(alpha))
    cos_b = 0.0 if beta == 90.0 else np.cos(np.deg2rad(beta))
    if gamma == 90.0:
        cos_g, sin_g = 0.0, 1.0
    else:
        g = np.deg2rad(gamma)
        cos_g, sin_g = np.cos(g), np.sin(g)
    box[0, 0] = lx
    box[1, 0] = ly * cos_g
    box[1, 1] = ly * sin_g
    box[2, 0] = lz * cos_b
    box[2, 1] = lz * (cos_a - cos_b * cos_g) / sin_g
    box[2, 2] = np.sqrt(lz * lz - np.float64(box[2, 0]) ** 2 - np.float64(box[2, 1]) ** 2)
    return box


def _wrap_triclinic(coords, box):
    """Move float32 coordinates into the primary triclinic cell: c axis, then b axis, then a
    axis, each against the slanted bounds of the parallelepiped (`_triclinic_pbc`)."""
    out = np.array(coords, dtype=np.float32, copy=True)
    b = box.astype(np.float64)
    b_z = b[2, 1] / b[2, 2]
    a_y = b[1, 0] / b[1, 1]
    a_z = (b[1, 1] * b[2, 0] - b[1, 0] * b[2, 1]) / (b[1, 1] * b[2, 2])
    for i in range(len(out)):
        c = out[i].astype(np.float64)
        s = np.floor(c[2] / b[2, 2])
        c -= s * b[2]
        s = np.floor((c[1] - c[2] * b_z) / b[1, 1])
        c -= s * b[1]
        s = np.floor((c[0] - (c[1] * a_y + c[2] * a_z)) / b[0, 0])
        c -= s * b[0]
        out[i] = c.astype(np.float32)
    return out


def _distances(ref, conf, box6):
    ref = np.asarray(ref, dtype=np.float32).reshape(-1, 3)
    conf = np.asarray(conf, dtype=np.float32).reshape(-1, 3)
    if box6 is None:
        d = (conf[None, :, :] - ref[:, None, :]).astype(np.float64)
        return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
    box6 = np.asarray(box6, dtype=np.float32)
    if np.all(box6[3:] == 90.0):
        box = box6[:3]
        inv = (1.0 / box.astype(np.float64)).astype(np.float32)
        dx = (conf[None, :, :] - ref[:, None, :]).astype(np.float64)  # f32 diff
        s = inv.astype(np.float64) * dx
        dx = box.astype(np.float64) * (s - np.round(s))
        return np.sqrt(dx[..., 0] * dx[..., 0] + dx[..., 1] * dx[..., 1] + dx[..., 2] * dx[..., 2])
    m = triclinic_vectors(box6)
    ref_w = _wrap_triclinic(ref, m)
    conf_w = _wrap_triclinic(conf, m)
    mb = m.astype(np.float64)
    out = np.empty((len(ref_w), len(conf_w)), dtype=np.float64)
    for i in range(len(ref_w)):
        dx = (conf_w - ref_w[i]).astype(np.float64)
        best = np.full(len(conf_w), np.finfo(np.float32).max, dtype=np.float64)
        for ix in (-1, 0, 1):
            rx = dx[:, 0] + mb[0, 0] * ix
            for iy in (-1, 0, 1):
                ry0 = rx + mb[1, 0] * iy
                ry1 = dx[:, 1] + mb[1, 1] * iy
                for iz in (-1, 0, 1):
                    rz0 = ry0 + mb[2, 0] * iz
                    rz1 = ry1 + mb[2, 1] * iz
                    rz2 = dx[:, 2] + mb[2, 2] * iz
                    dsq = rz0 * rz0 + rz1 * rz1 + rz2 * rz2
                    best = np.where(dsq < best, dsq, best)
        out[i] = np.sqrt(best)
    return out


def capped_distance(
    reference,
    configuration,
    max_cutoff,
    min_cutoff=None,
    box=None,
    method=None,
    return_distances=True,
):
    dist = _distances(reference, configuration, box)
    mask = dist <= max_cutoff
    if min_cutoff is not None:
        mask &= dist > min_cutoff
    ii, jj = np.nonzero(mask)
    pairs = np.c_[ii, jj].astype(np.intp)
    if return_distances:
        return pairs, dist[ii, jj]
    return pairs
