"""Synthetic reduced 3D-1D EMI system (BASELINE.json configs[4]).

The reference assembles it on a neuron mesh that has to be downloaded (src/emi_3d1d.py:28-43,
downloads.sh:11); offline the 1-D mesh is a seeded random tree of UnitCube MESH EDGES, which keeps
the construction of the reference: the 1-D mesh is an EmbeddedMesh of marked edges
(src/emi_3d1d.py:52), so its vertices are 3-D mesh vertices and, for radius 0, the coupling operator
is the 3D->1D trace = a vertex selection (src/emi_3d1d.py:63-68).

    a00 = k3 (grad u, grad v) + k3 (u, v)          3-D, P1        src/emi_3d1d.py:79
    a11 = k1 (p', q') + k1 (p, q)                  1-D, P1        src/emi_3d1d.py:80
    coupling  gamma [[Pi' M Pi, -Pi' M], [-M Pi, M]], M = 1-D mass   src/emi_3d1d.py:82-86
    Pi = trace (radius 0) or circle average (radius > 0, `circle_average`)   src/emi_3d1d.py:62-68
    f3 = x + y, f1 = 1                                            src/emi_3d1d.py:75
    interface dofs = every 1-D dof (src/utils.py:321); no Dirichlet conditions (pure Neumann + mass)
"""
import numpy as np
import scipy.sparse as sp

from . import Space, System, scalar_p1

# edge directions of the Kuhn triangulation (a vertex's 14 neighbours)
_DIRS = np.array([(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)])


def segment_graph(n, nsegments=4000, seed=0, branch_prob=0.02):
    """Seeded random tree on the mesh edges: returns (vertex ids [m], edges [nseg, 2] into that list)."""
    rng = np.random.default_rng(seed)
    start = np.array([n // 2, n // 2, n // 2])
    verts = {tuple(start): 0}
    coords = [start]
    edges = []
    tips = [0]
    dirs = np.concatenate([_DIRS, -_DIRS])
    guard = 0
    while len(edges) < nsegments and guard < 50 * nsegments:
        guard += 1
        t = tips[rng.integers(len(tips))] if rng.random() < branch_prob or len(tips) == 1 else tips[-1]
        d = dirs[rng.integers(len(dirs))]
        nxt = coords[t] + d
        if np.any(nxt < 0) or np.any(nxt > n) or tuple(nxt) in verts:
            if rng.random() < 0.3:
                tips.append(rng.integers(len(coords)))   # stuck: restart from a random node of the tree
            continue
        verts[tuple(nxt)] = len(coords)
        coords.append(nxt)
        edges.append((t, len(coords) - 1))
        tips.append(len(coords) - 1)
        if len(tips) > 64:
            tips = tips[-64:]
    coords = np.array(coords)
    ids = coords[:, 0] + (n + 1) * (coords[:, 1] + (n + 1) * coords[:, 2])
    return ids.astype(np.int64), np.array(edges, dtype=np.int64), coords / n


def circle_average(n, xyz1, edges, radius, npoints=12):
    """Pi (n1 x n3): nodal restatement of fenics_ii's Average(u, meshQ, Circle(radius, degree))
    (src/emi_3d1d.py:63-66): the mean of the P1 function u over the circle of the given radius around
    every 1-D vertex, in the plane normal to the curve there (tangent = mean direction of the incident
    segments).  The circle is sampled at `npoints` equispaced angles (the trapezoid rule is spectrally
    accurate on a circle); every sample is interpolated in the Kuhn tetrahedron that contains it, so a
    row of Pi couples the 1-D vertex to the 3-D dofs of all tetrahedra its circle crosses.  Samples
    outside the unit cube are clamped onto it."""
    n1 = len(xyz1)
    t = np.zeros((n1, 3))
    d = xyz1[edges[:, 1]] - xyz1[edges[:, 0]]
    d /= np.linalg.norm(d, axis=1)[:, None]
    np.add.at(t, edges[:, 0], d)
    np.add.at(t, edges[:, 1], d)
    nt = np.linalg.norm(t, axis=1)
    t[nt < 1e-12] = (1.0, 0.0, 0.0)      # a cusp whose directions cancel: any plane
    t /= np.linalg.norm(t, axis=1)[:, None]
    # orthonormal frame: cross with the coordinate axis least aligned with t
    ax = np.eye(3)[np.argmin(np.abs(t), axis=1)]
    e1 = np.cross(t, ax)
    e1 /= np.linalg.norm(e1, axis=1)[:, None]
    e2 = np.cross(t, e1)
    th = 2.0 * np.pi * np.arange(npoints) / npoints
    pts = (xyz1[:, None, :] + radius * (np.cos(th)[None, :, None] * e1[:, None, :]
                                        + np.sin(th)[None, :, None] * e2[:, None, :])).reshape(-1, 3)
    pts = np.clip(pts, 0.0, 1.0)
    g = pts * n
    cell = np.minimum(np.floor(g).astype(np.int64), n - 1)
    xi = g - cell
    order = np.argsort(-xi, axis=1, kind="stable")            # Kuhn simplex: coordinates in descending order
    xs = np.take_along_axis(xi, order, axis=1)
    lam = np.stack([1.0 - xs[:, 0], xs[:, 0] - xs[:, 1], xs[:, 1] - xs[:, 2], xs[:, 2]], axis=1)
    v = cell.copy()
    stride = np.array([1, n + 1, (n + 1) ** 2])
    cols = [v @ stride]
    for k in range(3):
        v = v + np.eye(3, dtype=np.int64)[order[:, k]]
        cols.append(v @ stride)
    rows = np.repeat(np.arange(n1), npoints)
    Pi = sp.coo_matrix((lam.T.ravel() / npoints, (np.tile(rows, 4), np.concatenate(cols))), shape=(n1, (n + 1) ** 3))
    Pi = Pi.tocsr()
    Pi.sum_duplicates()
    Pi.data[np.abs(Pi.data) < 1e-15] = 0.0
    Pi.eliminate_zeros()
    return Pi


def emi3d1d_system(n=32, gamma=1.0, k3=3.0, k1=7.0 * np.pi, nsegments=None, seed=0, radius=0.0):
    """radius = 0: trace coupling (Pi = vertex selection); radius > 0: perimeter-averaged coupling."""
    if radius < 0.0:
        raise ValueError("radius must be >= 0")
    nseg = nsegments if nsegments is not None else max(16, 4 * n)
    ids, edges, xyz1 = segment_graph(n, nseg, seed)
    n3, n1 = (n + 1) ** 3, len(ids)
    A3 = scalar_p1(3, [n, n, n], [1.0 / n] * 3, cK=k3, cM=k3)
    M3 = scalar_p1(3, [n, n, n], [1.0 / n] * 3, cK=0.0, cM=1.0)
    # 1-D P1 on the edge graph
    L = np.linalg.norm(xyz1[edges[:, 0]] - xyz1[edges[:, 1]], axis=1)
    r = np.concatenate([edges[:, 0], edges[:, 0], edges[:, 1], edges[:, 1]])
    c = np.concatenate([edges[:, 0], edges[:, 1], edges[:, 0], edges[:, 1]])
    K1 = sp.coo_matrix((np.concatenate([1 / L, -1 / L, -1 / L, 1 / L]), (r, c)), shape=(n1, n1)).tocsr()
    M1 = sp.coo_matrix((np.concatenate([L / 3, L / 6, L / 6, L / 3]), (r, c)), shape=(n1, n1)).tocsr()
    if radius > 0.0:
        Pi = circle_average(n, xyz1, edges, radius)
    else:
        Pi = sp.csr_matrix((np.ones(n1), (np.arange(n1), ids)), shape=(n1, n3))
    A00 = A3 + gamma * (Pi.T @ M1 @ Pi)
    A01 = -gamma * (Pi.T @ M1)
    A11 = k1 * (K1 + M1) + gamma * M1
    A = sp.bmat([[A00, A01], [A01.T, A11]], format="csr")
    A.sort_indices()
    idx = np.indices((n + 1,) * 3)[::-1]
    x3 = np.stack([idx[a].ravel() / n for a in range(3)], axis=1)
    b = np.concatenate([M3 @ (x3[:, 0] + x3[:, 1]), M1 @ np.ones(n1)])
    W = [Space(n3, x3), Space(n1, xyz1)]
    idofs = np.arange(n3, n3 + n1, dtype=np.int32)   # src/utils.py:321
    s = System("emi_3d1d", A, W, idofs, np.zeros(0, np.int32), 3, n,
               dict(gamma=gamma, k3=k3, k1=k1, nsegments=len(edges), seed=seed, radius=radius))
    s.b = b
    return s
