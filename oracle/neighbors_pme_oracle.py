"""TEST INFRASTRUCTURE ONLY -- numpy restatements of the reference's getNeighborPairs and PME arithmetic.

getNeighborPairs: src/pytorch/neighbors/getNeighborPairsCPU.cpp:56-98 (triangular pair enumeration, row > col, sequential
minimum image by DIVISION in the order z, y, x with round-half-to-even, inclusive cutoff on the distance, -1 / NaN padding).
PME direct: src/pytorch/pme/pmeCPU.cpp:74-172; PME reciprocal: pmeCPU.cpp:27-72 (spline), :174-277 (spread, rFFT, convolution,
energy), :279-353 (irFFT, interpolation); moduli and self energy: src/pytorch/pme/pme.py:94-129,194.

Parity pin: tests/test_oracle_neighbors_pme.py checks these against the reference's doctest (getNeighborPairs.py:104-138), the
np.tril_indices construction of TestNeighbors.py, and the OpenMM energies/forces in tests/golden/pme_openmm.json
(transcribed from TestPme.py:22-171).  O(N^2): small cases only.
"""
import math

import numpy as np
from scipy.special import erf, erfc


def neighbor_pairs(positions, cutoff, max_num_pairs=-1, box=None):
    """-> (neighbors int32 [2, P], deltas [P, 3], distances [P], num_found) in the dtype of `positions`."""
    pos = np.asarray(positions)
    dt = pos.dtype
    n = pos.shape[0]
    rows, cols = np.tril_indices(n, -1)          # row-major enumeration of row > col == the triangular index order
    deltas = (pos[rows] - pos[cols]).astype(dt)
    if box is not None:
        b = np.asarray(box, dt)
        for k in (2, 1, 0):
            s = np.rint((deltas[:, k] / b[k, k]).astype(dt)).astype(dt)
            deltas = (deltas - np.outer(s, b[k]).astype(dt)).astype(dt)
    d2 = (deltas[:, 0] * deltas[:, 0]).astype(dt)
    d2 = (d2 + (deltas[:, 1] * deltas[:, 1]).astype(dt)).astype(dt)
    d2 = (d2 + (deltas[:, 2] * deltas[:, 2]).astype(dt)).astype(dt)
    dist = np.sqrt(d2).astype(dt)
    keep = dist <= dt.type(cutoff)
    found = int(keep.sum())
    if max_num_pairs == -1:
        nb = np.stack([rows, cols]).astype(np.int32)
        nb[:, ~keep] = -1
        deltas = deltas.copy(); dist = dist.copy()
        deltas[~keep] = np.nan; dist[~keep] = np.nan
        return nb, deltas, dist, found
    nb = np.stack([rows[keep], cols[keep]]).astype(np.int32)[:, :max_num_pairs]
    deltas, dist = deltas[keep][:max_num_pairs], dist[keep][:max_num_pairs]
    pad = max_num_pairs - nb.shape[1]
    if pad > 0:
        nb = np.hstack([nb, np.full((2, pad), -1, np.int32)])
        deltas = np.vstack([deltas, np.full((pad, 3), np.nan, dt)])
        dist = np.hstack([dist, np.full(pad, np.nan, dt)])
    return nb, deltas, dist, found


def neighbor_pairs_backward(neighbors, deltas, distances, grad_deltas, grad_distances, n):
    g = np.zeros((n, 3), np.float64)
    for i in range(neighbors.shape[1]):
        r, c = neighbors[0, i], neighbors[1, i]
        if r < 0:
            continue
        v = grad_deltas[i] + deltas[i] / distances[i] * grad_distances[i]
        g[r] += v; g[c] -= v
    return g


# ---------------------------------------------------------------------------------------------------------------- PME
def _recip_box(box):
    b = np.asarray(box, np.float64)
    s = 1.0 / (b[0, 0] * b[1, 1] * b[2, 2])
    r = np.zeros((3, 3))
    r[0, 0] = b[1, 1] * b[2, 2] * s
    r[1, 0] = -b[1, 0] * b[2, 2] * s; r[1, 1] = b[0, 0] * b[2, 2] * s
    r[2, 0] = (b[1, 0] * b[2, 1] - b[1, 1] * b[2, 0]) * s; r[2, 1] = -b[0, 0] * b[2, 1] * s; r[2, 2] = b[0, 0] * b[1, 1] * s
    return r


def _spline(p, box, recip, grid, order):
    p = np.array(p, np.float64)
    b = np.asarray(box, np.float64)
    for i in (2, 1, 0):
        p = p - math.floor(p[i] * recip[i, i]) * b[i]
    t = p @ recip
    t = (t - np.floor(t)) * np.asarray(grid)
    ti = t.astype(np.int64)
    dr = t - ti
    gi = ti % np.asarray(grid)
    data = np.zeros((order, 3)); ddata = np.zeros((order, 3))
    for a in range(3):
        d = np.zeros(order)
        d[0], d[1] = 1 - dr[a], dr[a]
        for j in range(3, order + 1):
            if j == order:
                ddata[0, a] = -d[0]
                for k in range(1, order):
                    ddata[k, a] = d[k - 1] - d[k]
            new = np.zeros(order)
            for m in range(j):
                lo = d[m - 1] if m > 0 else 0.0
                hi = d[m] if m < j - 1 else 0.0
                new[m] = ((dr[a] + (j - 1 - m)) * lo + ((m + 1) - dr[a]) * hi) / (j - 1)
            d = new
        data[:, a] = d
    return gi, data, ddata


def pme_moduli(order, sizes):
    data = np.zeros(order)
    data[0] = 1
    for i in range(3, order):
        data[i - 1] = 0
        for j in range(1, i - 1):
            data[i - j - 1] = (j * data[i - j - 2] + (i - j) * data[i - j - 1]) / (i - 1)
        data[0] /= i - 1
    for i in range(1, order - 1):
        data[order - i - 1] = (i * data[order - i - 2] + (order - i) * data[order - i - 1]) / (order - 1)
    data[0] /= order - 1
    out = []
    for n in sizes:
        bs = np.zeros(max(n, order + 1)); bs[1:order + 1] = data; bs = bs[:n]
        k = np.arange(n)
        arg = 2 * np.pi * np.outer(k, k) / n
        m = (bs * np.cos(arg)).sum(1) ** 2 + (bs * np.sin(arg)).sum(1) ** 2
        for i in range(n):
            if m[i] < 1e-7:
                m[i] = (m[(i - 1) % n] + m[(i + 1) % n]) * 0.5
        out.append(m)
    return out


def pme_reciprocal(pos, charges, box, grid, order, alpha, coulomb):
    """-> (energy incl. self term, dE/dpos [N,3], dE/dq [N]) in float64."""
    pos = np.asarray(pos, np.float64); q = np.asarray(charges, np.float64); box = np.asarray(box, np.float64)
    gx, gy, gz = grid
    recip = _recip_box(box)
    sc = math.sqrt(coulomb)
    Q = np.zeros(grid)
    splines = [_spline(p, box, recip, grid, order) for p in pos]
    for (gi, data, _), qi in zip(splines, q):
        for ix in range(order):
            for iy in range(order):
                for iz in range(order):
                    Q[(gi[0] + ix) % gx, (gi[1] + iy) % gy, (gi[2] + iz) % gz] += qi * sc * data[ix, 0] * data[iy, 1] * data[iz, 2]
    F = np.fft.rfftn(Q)
    xm, ym, zm = pme_moduli(order, grid)
    energy = 0.0
    scale_factor = math.pi * box[0, 0] * box[1, 1] * box[2, 2]
    exp_factor = math.pi ** 2 / alpha ** 2
    for kx in range(gx):
        mx = kx if kx < (gx + 1) // 2 else kx - gx
        for ky in range(gy):
            my = ky if ky < (gy + 1) // 2 else ky - gy
            for kz in range(gz // 2 + 1):
                mz = kz if kz < (gz + 1) // 2 else kz - gz
                mh = np.array([mx * recip[0, 0], mx * recip[1, 0] + my * recip[1, 1], mx * recip[2, 0] + my * recip[2, 1] + mz * recip[2, 2]])
                m2 = mh @ mh
                if kx == 0 and ky == 0 and kz == 0:
                    e = 0.0
                else:
                    e = math.exp(-exp_factor * m2) / (m2 * scale_factor * xm[kx] * ym[ky] * zm[kz])
                w = 2.0 if 0 < kz <= (gz - 1) // 2 else 1.0
                energy += w * e * abs(F[kx, ky, kz]) ** 2
                F[kx, ky, kz] *= e
    phi = np.fft.irfftn(F, s=grid, axes=(0, 1, 2), norm="forward")
    dpos = np.zeros((len(pos), 3)); dq = np.zeros(len(pos))
    for a, ((gi, data, ddata), qi) in enumerate(zip(splines, q)):
        d = np.zeros(3); s = 0.0
        for ix in range(order):
            for iy in range(order):
                for iz in range(order):
                    g = phi[(gi[0] + ix) % gx, (gi[1] + iy) % gy, (gi[2] + iz) % gz]
                    d[0] += ddata[ix, 0] * data[iy, 1] * data[iz, 2] * g
                    d[1] += data[ix, 0] * ddata[iy, 1] * data[iz, 2] * g
                    d[2] += data[ix, 0] * data[iy, 1] * ddata[iz, 2] * g
                    s += data[ix, 0] * data[iy, 1] * data[iz, 2] * g
        dpos[a, 0] = qi * sc * (d[0] * gx * recip[0, 0])
        dpos[a, 1] = qi * sc * (d[0] * gx * recip[1, 0] + d[1] * gy * recip[1, 1])
        dpos[a, 2] = qi * sc * (d[0] * gx * recip[2, 0] + d[1] * gy * recip[2, 1] + d[2] * gz * recip[2, 2])
        dq[a] = s * sc
    self_energy = -np.sum(q ** 2) * coulomb * alpha / math.sqrt(math.pi)
    dq_self = -2 * q * coulomb * alpha / math.sqrt(math.pi)
    return 0.5 * energy + self_energy, dpos, dq + dq_self


def pme_direct(pos, charges, box, cutoff, alpha, coulomb, exclusions=None):
    """-> (energy, dE/dpos, dE/dq) in float64; pair list from neighbor_pairs (float64)."""
    pos = np.asarray(pos, np.float64); q = np.asarray(charges, np.float64)
    nb, deltas, dist, _ = neighbor_pairs(pos, cutoff, -1, np.asarray(box, np.float64))
    n = len(pos)
    excl = [set() for _ in range(n)]
    if exclusions is not None:
        for i, row in enumerate(np.asarray(exclusions)):
            excl[i] = {int(x) for x in row if x >= 0}
    energy = 0.0; dpos = np.zeros((n, 3)); dq = np.zeros(n)
    for i in range(nb.shape[1]):
        a1, a2 = nb[0, i], nb[1, i]
        if a1 < 0 or a2 in excl[a1]:
            continue
        r = dist[i]; ar = alpha * r; pref = coulomb / r
        energy += pref * erfc(ar) * q[a1] * q[a2]
        dq[a1] += pref * erfc(ar) * q[a2]; dq[a2] += pref * erfc(ar) * q[a1]
        dedr = pref * q[a1] * q[a2] * (erfc(ar) + ar * math.exp(-ar * ar) * 2 / math.sqrt(math.pi)) / r ** 2
        dpos[a1] -= dedr * deltas[i]; dpos[a2] += dedr * deltas[i]
    for a1 in range(n):
        for a2 in excl[a1]:
            if a2 <= a1:
                continue
            dr = pos[a1] - pos[a2]; r = math.sqrt(dr @ dr); ar = alpha * r; pref = coulomb / r
            energy -= pref * erf(ar) * q[a1] * q[a2]
            dq[a1] -= pref * erf(ar) * q[a2]; dq[a2] -= pref * erf(ar) * q[a1]
            dedr = pref * q[a1] * q[a2] * (erf(ar) - ar * math.exp(-ar * ar) * 2 / math.sqrt(math.pi)) / r ** 2
            dpos[a1] += dedr * dr; dpos[a2] -= dedr * dr
    return energy, dpos, dq


# -------------------------------------------------------------------------------------------- vectorised PME restatement
# Same arithmetic as _spline / pme_reciprocal above (src/pytorch/pme/pmeCPU.cpp:197-349), written with whole-array numpy
# operations so that BASELINE config 5 (200 000 charges, 128^3 grid, order 5) finishes in seconds.  tests/ pins it to the loop
# version on small systems.
def _spline_vec(pos, box, recip, grid, order):
    p = np.array(pos, np.float64)
    b = np.asarray(box, np.float64)
    for i in (2, 1, 0):
        p = p - np.floor(p[:, i] * recip[i, i])[:, None] * b[i][None, :]
    t = p @ recip
    t = (t - np.floor(t)) * np.asarray(grid)
    ti = t.astype(np.int64)
    dr = t - ti
    gi = ti % np.asarray(grid)
    n = len(p)
    data = np.zeros((n, order, 3)); ddata = np.zeros((n, order, 3))
    d = np.zeros((n, order, 3))
    d[:, 0], d[:, 1] = 1 - dr, dr
    for j in range(3, order + 1):
        if j == order:
            ddata[:, 0] = -d[:, 0]
            for k in range(1, order):
                ddata[:, k] = d[:, k - 1] - d[:, k]
        new = np.zeros_like(d)
        for m in range(j):
            lo = d[:, m - 1] if m > 0 else 0.0
            hi = d[:, m] if m < j - 1 else 0.0
            new[:, m] = ((dr + (j - 1 - m)) * lo + ((m + 1) - dr) * hi) / (j - 1)
        d = new
    data[:] = d
    return gi, data, ddata


def pme_reciprocal_vec(pos, charges, box, grid, order, alpha, coulomb):
    """Vectorised pme_reciprocal: -> (energy incl. self term, dE/dpos [N,3], dE/dq [N]) in float64."""
    pos = np.asarray(pos, np.float64); q = np.asarray(charges, np.float64); box = np.asarray(box, np.float64)
    gx, gy, gz = grid
    recip = _recip_box(box)
    sc = math.sqrt(coulomb)
    gi, data, ddata = _spline_vec(pos, box, recip, grid, order)
    o = np.arange(order)
    ix = (gi[:, 0, None] + o) % gx; iy = (gi[:, 1, None] + o) % gy; iz = (gi[:, 2, None] + o) % gz          # [N, order]
    flat = (ix[:, :, None, None] * gy + iy[:, None, :, None]) * gz + iz[:, None, None, :]                      # [N, o, o, o]
    w = data[:, :, 0][:, :, None, None] * data[:, :, 1][:, None, :, None] * data[:, :, 2][:, None, None, :]
    Q = np.bincount(flat.ravel(), weights=(w * (q * sc)[:, None, None, None]).ravel(), minlength=gx * gy * gz).reshape(grid)
    F = np.fft.rfftn(Q)
    xm, ym, zm = pme_moduli(order, grid)
    kx = np.arange(gx); ky = np.arange(gy); kz = np.arange(gz // 2 + 1)
    mx = np.where(kx < (gx + 1) // 2, kx, kx - gx)[:, None, None].astype(np.float64)
    my = np.where(ky < (gy + 1) // 2, ky, ky - gy)[None, :, None].astype(np.float64)
    mz = np.where(kz < (gz + 1) // 2, kz, kz - gz)[None, None, :].astype(np.float64)
    hx = mx * recip[0, 0]; hy = mx * recip[1, 0] + my * recip[1, 1]; hz = mx * recip[2, 0] + my * recip[2, 1] + mz * recip[2, 2]
    m2 = hx * hx + hy * hy + hz * hz
    scale_factor = math.pi * box[0, 0] * box[1, 1] * box[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.exp(-(math.pi ** 2 / alpha ** 2) * m2) / (m2 * scale_factor * np.asarray(xm)[:, None, None] * np.asarray(ym)[None, :, None] *
                                                        np.asarray(zm)[None, None, : gz // 2 + 1])
    e[0, 0, 0] = 0.0
    wk = np.where((kz > 0) & (kz <= (gz - 1) // 2), 2.0, 1.0)[None, None, :]
    energy = float(np.sum(wk * e * np.abs(F) ** 2))
    phi = np.fft.irfftn(F * e, s=grid, axes=(0, 1, 2), norm="forward").ravel()
    g = phi[flat]                                                                                                # [N, o, o, o]
    d0 = np.einsum("nxyz,nx,ny,nz->n", g, ddata[:, :, 0], data[:, :, 1], data[:, :, 2])
    d1 = np.einsum("nxyz,nx,ny,nz->n", g, data[:, :, 0], ddata[:, :, 1], data[:, :, 2])
    d2 = np.einsum("nxyz,nx,ny,nz->n", g, data[:, :, 0], data[:, :, 1], ddata[:, :, 2])
    s = np.einsum("nxyz,nx,ny,nz->n", g, data[:, :, 0], data[:, :, 1], data[:, :, 2])
    dpos = np.stack([q * sc * (d0 * gx * recip[0, 0]),
                     q * sc * (d0 * gx * recip[1, 0] + d1 * gy * recip[1, 1]),
                     q * sc * (d0 * gx * recip[2, 0] + d1 * gy * recip[2, 1] + d2 * gz * recip[2, 2])], axis=1)
    self_energy = -np.sum(q ** 2) * coulomb * alpha / math.sqrt(math.pi)
    dq_self = -2 * q * coulomb * alpha / math.sqrt(math.pi)
    return 0.5 * energy + self_energy, dpos, s * sc + dq_self


def pme_direct_sampled(pos, charges, box_edge, cutoff, alpha, coulomb, sample):
    """Direct-space PME of a CUBIC periodic box restricted to the atoms `sample`: -> (dE/dpos [len(sample), 3], dE/dq [len(sample)],
    per-atom energy share) in float64, neighbours from a periodic KD-tree (pmeCPU.cpp:105-157, no exclusions).  For systems too
    large for the all-pairs loop of pme_direct."""
    from scipy.spatial import cKDTree
    from scipy.special import erfc as erfc_v
    pos = np.asarray(pos, np.float64); q = np.asarray(charges, np.float64)
    L = float(box_edge)
    w = np.mod(pos, L)
    tree = cKDTree(w, boxsize=L)
    dpos = np.zeros((len(sample), 3)); dq = np.zeros(len(sample)); eshare = np.zeros(len(sample))
    for k, i in enumerate(sample):
        nb = np.array([j for j in tree.query_ball_point(w[i], cutoff) if j != i], np.int64)
        d = w[i] - w[nb]
        d -= np.round(d / L) * L                      # delta = pos[i] - pos[j], minimum image
        r = np.sqrt((d * d).sum(1))
        keep = r < cutoff
        nb, d, r = nb[keep], d[keep], r[keep]
        ar = alpha * r; pref = coulomb / r
        eshare[k] = 0.5 * np.sum(pref * erfc_v(ar) * q[i] * q[nb])
        dq[k] = np.sum(pref * erfc_v(ar) * q[nb])
        dedr = pref * q[i] * q[nb] * (erfc_v(ar) + ar * np.exp(-ar * ar) * 2 / math.sqrt(math.pi)) / r ** 2
        dpos[k] = -(dedr[:, None] * d).sum(0)
    return dpos, dq, eshare
