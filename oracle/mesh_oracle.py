"""CPU oracle for the iso-surface extraction step -- TEST INFRASTRUCTURE ONLY.  ** parity unpinned **

Reference: oai_analysis/mesh_processing.py:325-340 (get_mesh: skimage.measure.marching_cubes(level=0.5, spacing,
step_size=1, gradient_direction="ascent") on the atlas-space probability map, axes swapped to x,y,z) and :102-146
(get_vtk_mesh: vtkPolyDataConnectivityFilter regions, those with more than 3000 cells are kept).

skimage and vtk are not installable offline and their sources are not under /root/reference, so this file restates the
published algorithm instead of calling it:
  * marching cubes with vertices on the lattice edges the iso-level crosses, linearly interpolated (what skimage's
    Lewiner implementation computes for an edge vertex), ONE vertex per crossed edge (shared between cells);
  * ambiguous cube faces (two diagonal corners inside) are resolved with the asymptotic decider (the bilinear saddle
    value against the level) -- the face test of Lewiner et al. 2003 / Chernyaev's MC33 -- so the surface is watertight
    and consistent between neighbouring cells; the INTERIOR tests of MC33 (tunnel cases, the extra centre vertex of
    case 13.5) are not restated, so on the rare cells where they matter skimage may triangulate the inside of the cell
    differently (same vertices on the cell's edges, same boundary loops on its faces);
  * triangle winding: normals point towards ascending values for gradient_direction="ascent".
  * regions = connected components of the triangles over shared vertices; kept when they have MORE than 3000 cells.
Nothing in the reference's tests pins a vertex or face count offline (test/test_all.py:69-70 are commented out), hence
"parity unpinned".  The polygon table below is generated, not copied: for every corner-sign pattern and every
resolution of its ambiguous faces the crossed edges are chained into closed loops face by face.
"""
import numpy as np

# corner index = dx + 2 dy + 4 dz; edge id = 4 * axis + (u + 2 v) with (u, v) the offsets along the two other axes in
# increasing axis order; an edge runs from its base corner to base + 1 along `axis`
CORNERS = np.array([[(c >> 0) & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)])


def _edge_corners(e):
    axis, uv = e // 4, e % 4
    others = [a for a in range(3) if a != axis]
    base = [0, 0, 0]
    base[others[0]], base[others[1]] = uv & 1, uv >> 1
    c0 = base[0] + 2 * base[1] + 4 * base[2]
    return c0, c0 + (1 << axis)


EDGE_CORNERS = [_edge_corners(e) for e in range(12)]
EDGE_MID = np.array([(CORNERS[a] + CORNERS[b]) / 2.0 for a, b in EDGE_CORNERS])


def _edge_between(c0, c1):
    for e, (a, b) in enumerate(EDGE_CORNERS):
        if {a, b} == {c0, c1}:
            return e
    raise KeyError((c0, c1))


def _faces():
    """6 faces (axis, side): corners in cyclic order, outward normal."""
    out = []
    for axis in range(3):
        o = [a for a in range(3) if a != axis]
        for side in range(2):
            cyc = []
            for u, v in ((0, 0), (1, 0), (1, 1), (0, 1)):
                p = [0, 0, 0]
                p[axis], p[o[0]], p[o[1]] = side, u, v
                cyc.append(p[0] + 2 * p[1] + 4 * p[2])
            n = np.zeros(3)
            n[axis] = 1.0 if side else -1.0
            out.append((cyc, n))
    return out


FACES = _faces()


def face_is_ambiguous(mask, f):
    cyc = FACES[f][0]
    s = [(mask >> c) & 1 for c in cyc]
    return s[0] == s[2] and s[1] == s[3] and s[0] != s[1]


def cell_polygons(mask, connect_bits):
    """Closed, oriented loops of edge ids for the corner-inside mask; bit f of connect_bits says the INSIDE corners of
    ambiguous face f are connected across the face (saddle inside).  Loops are oriented so that their normal (right-hand
    rule) points from the inside corners to the outside ("descent"); callers flip for "ascent"."""
    nxt = {}
    for f, (cyc, nrm) in enumerate(FACES):
        s = [(mask >> c) & 1 for c in cyc]
        k = sum(s)
        if k == 0 or k == 4:
            continue
        segs = []   # (edge_a, edge_b, reference corner, reference corner is on the inside side)
        e = [_edge_between(cyc[i], cyc[(i + 1) % 4]) for i in range(4)]   # e[i] between corner i and i+1
        if k == 1 or k == 3:
            odd = s.index(1 if k == 1 else 0)
            segs.append((e[(odd - 1) % 4], e[odd], cyc[odd], k == 1))
        elif s[0] == s[1] or s[1] == s[2]:        # two adjacent corners inside
            i = [j for j in range(4) if s[j] and s[(j + 1) % 4]][0]
            segs.append((e[(i - 1) % 4], e[(i + 1) % 4], cyc[i], True))
        else:                                      # ambiguous: two diagonal corners inside
            cut_inside = not ((connect_bits >> f) & 1)
            for i in range(4):
                if bool(s[i]) == cut_inside:       # the segment cuts this corner off
                    segs.append((e[(i - 1) % 4], e[i], cyc[i], cut_inside))
        for ea, eb, ref, ref_inside in segs:
            a, b, p = EDGE_MID[ea], EDGE_MID[eb], CORNERS[ref].astype(float)
            left = float(np.dot(nrm, np.cross(b - a, p - a)))   # > 0: the reference corner is left of a -> b seen from outside
            # orientation rule: walking a -> b as seen from outside the cube, the inside region lies on the RIGHT, which
            # makes the loop's right-hand normal point away from the inside corners
            forward = (left < 0) == ref_inside
            src, dst = (ea, eb) if forward else (eb, ea)
            assert src not in nxt
            nxt[src] = dst
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, cur = [], start
        while cur not in seen:
            seen.add(cur)
            loop.append(cur)
            cur = nxt[cur]
        assert cur == start, "open loop"
        loops.append(loop)
    return loops


FACE_EDGES = [frozenset(_edge_between(cyc[i], cyc[(i + 1) % 4]) for i in range(4)) for cyc, _ in FACES]


def _coplanar(e0, e1):
    """both crossed edges lie on one cube face: a diagonal between their vertices would lie IN that face"""
    return any(e0 in fe and e1 in fe for fe in FACE_EDGES)


def cell_triangles(mask, connect_bits):
    """Fan triangulation of every loop: list of (e0, e1, e2) edge-id triples.  The apex is the first loop vertex (in
    loop order from the smallest edge id) whose fan has no diagonal lying in a cube face -- such a diagonal could
    coincide with the neighbouring cell's and make the edge non-manifold; loops where every apex has one keep apex 0."""
    tris = []
    for loop in cell_polygons(mask, connect_bits):
        n, best = len(loop), 0
        for apex in range(n):
            diag = [loop[(apex + k) % n] for k in range(2, n - 1)]
            if not any(_coplanar(loop[apex], d) for d in diag):
                best = apex
                break
        rot = loop[best:] + loop[:best]
        for i in range(1, n - 1):
            tris.append((rot[0], rot[i], rot[i + 1]))
    return tris


_TABLE = {}


def table_entry(mask, connect_bits):
    # only the bits of ambiguous faces matter
    key_bits = 0
    for f in range(6):
        if face_is_ambiguous(mask, f) and (connect_bits >> f) & 1:
            key_bits |= 1 << f
    key = (mask, key_bits)
    if key not in _TABLE:
        _TABLE[key] = cell_triangles(mask, key_bits)
    return _TABLE[key]


def marching_cubes(volume, level=0.5, spacing=(1.0, 1.0, 1.0), gradient_direction="ascent"):
    """volume indexed [i0, i1, i2]; vertex coordinates are index * spacing per axis (no origin), like skimage.
    Returns (verts float64 [n,3], faces int64 [m,3]).  Vertices are unique per crossed lattice edge."""
    v = np.asarray(volume, dtype=np.float64) - level
    n0, n1, n2 = v.shape
    inside = v > 0   # strictly above the level is inside (a value equal to the level is outside)
    sp = np.asarray(spacing, dtype=np.float64)
    vert_id, verts, faces = {}, [], []

    def vertex(i, j, k, axis):
        key = (i, j, k, axis)
        if key not in vert_id:
            p0 = np.array([i, j, k], dtype=np.float64)
            q = [i, j, k]
            q[axis] += 1
            a, b = v[i, j, k], v[q[0], q[1], q[2]]
            t = a / (a - b)
            p = p0.copy()
            p[axis] += t
            vert_id[key] = len(verts)
            verts.append(p * sp)
        return vert_id[key]

    # candidate cells: any sign change among the 8 corners
    c = inside
    any_in = (c[:-1, :-1, :-1] | c[1:, :-1, :-1] | c[:-1, 1:, :-1] | c[1:, 1:, :-1] |
              c[:-1, :-1, 1:] | c[1:, :-1, 1:] | c[:-1, 1:, 1:] | c[1:, 1:, 1:])
    all_in = (c[:-1, :-1, :-1] & c[1:, :-1, :-1] & c[:-1, 1:, :-1] & c[1:, 1:, :-1] &
              c[:-1, :-1, 1:] & c[1:, :-1, 1:] & c[:-1, 1:, 1:] & c[1:, 1:, 1:])
    for i, j, k in zip(*np.nonzero(any_in & ~all_in)):
        vals = [v[i + CORNERS[cc][0], j + CORNERS[cc][1], k + CORNERS[cc][2]] for cc in range(8)]
        mask = sum(1 << cc for cc in range(8) if vals[cc] > 0)
        bits = 0
        for f, (cyc, _) in enumerate(FACES):
            if face_is_ambiguous(mask, f):
                a, b, cc2, d = (vals[q] for q in cyc)
                # asymptotic decider: the bilinear saddle value (a c - b d) / (a + c - b - d), relative to the level
                saddle = (a * cc2 - b * d) / (a + cc2 - b - d)
                if saddle > 0:
                    bits |= 1 << f
        for tri in table_entry(mask, bits):
            ids = []
            for e in tri:
                c0 = EDGE_CORNERS[e][0]
                ids.append(vertex(i + CORNERS[c0][0], j + CORNERS[c0][1], k + CORNERS[c0][2], e // 4))
            if gradient_direction == "ascent":
                ids = ids[::-1]
            faces.append(ids)
    return np.asarray(verts, dtype=np.float64).reshape(-1, 3), np.asarray(faces, dtype=np.int64).reshape(-1, 3)


def keep_large_regions(verts, faces, min_cells=3000):
    """get_vtk_mesh (mesh_processing.py:119-141): connected regions (triangles sharing points) with more than
    `min_cells` cells are kept; vertices are compacted to the ones still referenced."""
    parent = np.arange(len(verts))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b, c in faces:
        ra, rb, rc = find(a), find(b), find(c)
        r = min(ra, rb, rc)
        parent[ra] = parent[rb] = parent[rc] = r
    root = np.array([find(x) for x in range(len(verts))])
    face_root = root[faces[:, 0]] if len(faces) else np.zeros(0, dtype=np.int64)
    counts = np.bincount(face_root, minlength=len(verts)) if len(faces) else np.zeros(len(verts), dtype=np.int64)
    keep = counts[face_root] > min_cells
    kept = faces[keep]
    used = np.zeros(len(verts), dtype=bool)
    used[kept.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    return verts[used], remap[kept], int(np.unique(face_root[keep]).size) if len(faces) else 0


def mesh_stats(verts, faces):
    """area, enclosed signed volume (divergence theorem), Euler characteristic V - E + F."""
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()
    vol = np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1)
    n_edges = len(np.unique(e, axis=0))
    return dict(area=float(area), volume=float(vol), euler=int(len(verts) - n_edges + len(faces)),
                n_verts=int(len(verts)), n_faces=int(len(faces)))


# ---------------------------------------------------------------------------------------------------------------
# Mesh post-processing (reference oai_analysis/mesh_processing.py:197-321) -- ** parity unpinned **
# except KMeans, which is the real sklearn call the reference makes.
# ---------------------------------------------------------------------------------------------------------------
def vertex_neighbours(n_verts, faces):
    """unique edge neighbours per vertex + flag of vertices on an open boundary edge (an edge used by one face)"""
    nbr = [set() for _ in range(n_verts)]
    edge_use = {}
    for a, b, c in faces:
        for u, v in ((a, b), (b, c), (c, a)):
            nbr[u].add(v)
            nbr[v].add(u)
            key = (min(u, v), max(u, v))
            edge_use[key] = edge_use.get(key, 0) + 1
    fixed = np.zeros(n_verts, dtype=bool)
    for (u, v), k in edge_use.items():
        if k == 1:
            fixed[u] = fixed[v] = True
    return [sorted(s) for s in nbr], fixed


def smooth_mesh(verts, faces, iterations=150, relaxation=0.01, in_place=True):
    """vtkSmoothPolyDataFilter with the defaults smooth_mesh leaves (mesh_processing.py:298-306): Laplacian relaxation
    towards the mean of the edge neighbours, points stored as float32 after every update.  in_place=True sweeps the
    vertices in index order reusing already-updated neighbours (VTK's loop); False is the Jacobi form the GPU runs.
    Open-boundary vertices are held fixed (closed iso-surfaces have none)."""
    nbr, fixed = vertex_neighbours(len(verts), faces)
    x = np.asarray(verts, dtype=np.float32).copy()
    for _ in range(iterations):
        src = x if in_place else x.copy()
        for i in range(len(x)):
            if fixed[i] or not nbr[i]:
                continue
            xi = src[i].astype(np.float64)
            d = np.zeros(3)
            for j in nbr[i]:
                d += (src[j].astype(np.float64) - xi) / len(nbr[i])
            x[i] = (xi + relaxation * d).astype(np.float32)
    return x


def face_normals_centroids(verts, faces):
    """trimesh face_normals (unit (b-a)x(c-a)) and get_cell_centroid (mesh_processing.py:25-47)."""
    v = np.asarray(verts, dtype=np.float64)
    a, b, c = v[faces[:, 0]], v[faces[:, 1]], v[faces[:, 2]]
    n = np.cross(b - a, c - a)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.divide(n, ln, out=np.zeros_like(n), where=ln > 0)
    return n, (a + b + c) / 3.0


def point_mesh_distance(points, verts, faces, chunk=256):
    """vtkDistancePolyDataFilter (unsigned): distance from each point to the closest point on any triangle."""
    v = np.asarray(verts, dtype=np.float64)
    A, B, C = v[faces[:, 0]], v[faces[:, 1]], v[faces[:, 2]]
    out = np.empty(len(points))
    P = np.asarray(points, dtype=np.float64)
    for s in range(0, len(P), chunk):
        p = P[s:s + chunk, None, :]
        ab, ac, ap = B - A, C - A, p - A
        d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
        bp = p - B
        d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
        cp = p - C
        d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
        va, vb, vc = d3 * d6 - d5 * d4, d5 * d2 - d1 * d6, d1 * d4 - d3 * d2
        with np.errstate(divide="ignore", invalid="ignore"):
            den = 1.0 / (va + vb + vc)
            q = A + ab * (vb * den)[..., None] + ac * (vc * den)[..., None]                    # interior
            t = (d4 - d3) / ((d4 - d3) + (d5 - d6))
            q = np.where(((va <= 0) & (d4 - d3 >= 0) & (d5 - d6 >= 0))[..., None], B + (C - B) * t[..., None], q)
            t = d2 / (d2 - d6)
            q = np.where(((vb <= 0) & (d2 >= 0) & (d6 <= 0))[..., None], A + ac * t[..., None], q)
            q = np.where(((d6 >= 0) & (d5 <= d6))[..., None], C, q)
            t = d1 / (d1 - d3)
            q = np.where(((vc <= 0) & (d1 >= 0) & (d3 <= 0))[..., None], A + ab * t[..., None], q)
            q = np.where(((d3 >= 0) & (d4 <= d3))[..., None], B, q)
            q = np.where(((d1 <= 0) & (d2 <= 0))[..., None], A, q)
        out[s:s + chunk] = np.sqrt(((p - q) ** 2).sum(-1).min(axis=1))
    return out


def split_tibial(normals, centroids):
    """split_tibial_cartilage_surface (mesh_processing.py:197-222) with the reference's own sklearn KMeans call."""
    from sklearn.cluster import KMeans
    cn = (centroids - centroids.mean(0)) / (centroids.max(0) - centroids.min(0))
    feats = np.concatenate((cn * 1, normals * 10), axis=1)
    labels = KMeans(n_clusters=2, algorithm="lloyd", random_state=5).fit(feats).labels_ * 2 - 1
    if normals[labels == -1, 1].mean() < 0:
        labels = -labels
    return labels, feats


def split_femoral(normals, centroids, bounds_min, bounds_max, num_divisions=3):
    """split_femoral_cartilage_surface (mesh_processing.py:243-294), sklearn KMeans(n_init=5, random_state=5)."""
    from sklearn.cluster import KMeans
    cn = (centroids - centroids.mean(0)) / (centroids.max(0) - centroids.min(0))
    center = (np.asarray(bounds_min) + np.asarray(bounds_max)) / 2
    dot = (center - centroids) * normals
    x = cn[:, 0]
    lo, step = x.min(), (x.max() - x.min()) / num_divisions
    out = np.zeros(len(cn))
    for i in range(num_divisions):
        idx = np.where((x >= lo + step * i) & (x < lo + step * i + step))[0]
        feats = np.concatenate((cn[idx], normals[idx], dot[idx]), axis=1)
        lab = KMeans(n_clusters=2, algorithm="lloyd", n_init=5, random_state=5).fit(feats).labels_ * 2 - 1
        if normals[idx][lab == -1, 1].mean() < 0:
            lab = -lab
        out[idx] = lab
    return out


# ------------------------------------------------------------------------------------------------ §8f-4
def map_attributes(source_pts, source_attr, target_pts, radius=1.0):
    """mesh_processing.py:398-406: vtkPointInterpolator(source -> target points) with its DEFAULT kernel and
    SetNullPointsStrategyToClosestPoint().  vtk is not installable here; restated from the VTK sources as published:
    the default kernel is vtkLinearKernel (a vtkGeneralizedKernel with footprint RADIUS, Radius = 1.0): the value at a
    target point is the plain average of the source attributes at the source points within `radius` (distance <=
    radius); a target point with no source point in range is a "null point" and takes the attribute of the closest
    source point.  ** parity unpinned ** (no reference test holds a mapped value).  float64, brute force."""
    s = np.asarray(source_pts, dtype=np.float64)
    a = np.asarray(source_attr, dtype=np.float64)
    t = np.asarray(target_pts, dtype=np.float64)
    a2 = a.reshape(len(s), -1)
    out = np.empty((len(t), a2.shape[1]))
    for i0 in range(0, len(t), 512):
        d2 = ((t[i0:i0 + 512, None, :] - s[None, :, :]) ** 2).sum(-1)
        inside = d2 <= radius * radius
        cnt = inside.sum(1)
        sums = inside.astype(np.float64) @ a2
        near = a2[d2.argmin(1)]
        out[i0:i0 + 512] = np.where(cnt[:, None] > 0, sums / np.maximum(cnt, 1)[:, None], near)
    return out.reshape((len(t),) + a.shape[1:])


def compute_least_square_circle(x, y):
    """mesh_processing.py:409-443 verbatim in behaviour: scipy.optimize.leastsq (the reference's own call; scipy is
    installed here, so this half IS the reference's arithmetic) on the algebraic distance to the mean circle, started
    at the centroid, with the analytic Jacobian."""
    from scipy import optimize
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)

    def calc_r(xc, yc):
        return np.sqrt((x - xc) ** 2 + (y - yc) ** 2)

    def f(c):
        r = calc_r(*c)
        return r - r.mean()

    def df(c):
        xc, yc = c
        r = calc_r(xc, yc)
        d = np.empty((2, x.size))
        d[0] = (xc - x) / r
        d[1] = (yc - y) / r
        return d - d.mean(axis=1)[:, None]

    center, _ = optimize.leastsq(f, (x.mean(), y.mean()), Dfun=df, col_deriv=True)
    return center, calc_r(*center).mean()


def project_thickness_fc(verts, thickness):
    """mesh_processing.py:489-496 (mesh_type == "FC"): swap x and y, fit the cylinder, unroll by the polar angle."""
    v = np.array(verts, dtype=np.float32)
    v[:, [1, 0]] = v[:, [0, 1]]
    center, _ = compute_least_square_circle(v[:, 0], v[:, 1])
    angle = np.arctan2(v[:, 1] - center[1], v[:, 0] - center[0])
    return angle, v[:, 2].astype(np.float64), np.asarray(thickness), center


def linear_kernel_pca2(points):
    """sklearn.decomposition.KernelPCA(n_components=2, degree=3.0) as the reference calls it (mesh_processing.py:472-
    477): the kernel defaults to "linear" (degree is ignored), so fit_transform returns the first two principal-component
    scores of the centred points, each column signed so that its entry of largest magnitude is positive (sklearn's
    svd_flip on the Gram matrix's eigenvectors).  Restated through the 3x3 covariance (an n x n Gram matrix of a 20 k
    vertex half-mesh does not fit a test); tests/test_mesh_oracle.py checks it against sklearn itself on small inputs."""
    x = np.asarray(points, dtype=np.float64)
    xc = x - x.mean(0)
    w, v = np.linalg.eigh(xc.T @ xc)
    order = np.argsort(w)[::-1][:2]
    scores = xc @ v[:, order]
    for k in range(2):
        j = np.argmax(np.abs(scores[:, k]))
        if scores[j, k] < 0:
            scores[:, k] = -scores[:, k]
    return scores


def _rotate_embedded(embedded, angle):
    theta = (angle / 180.0) * np.pi
    rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    return np.dot(embedded, rot)


def project_thickness_tc(verts, thickness):
    """mesh_processing.py:497-534 (tibial): split at z = 50, PCA-flatten each plateau, rotate (-50 / -160 degrees), mirror
    the right one, stack [right; left] with the right one lifted by 50."""
    v = np.asarray(verts)
    t = np.asarray(thickness)
    left, right = np.where(v[:, 2] < 50)[0], np.where(v[:, 2] >= 50)[0]
    el = _rotate_embedded(linear_kernel_pca2(v[left]), -50)
    er = _rotate_embedded(linear_kernel_pca2(v[right]), -160)
    er[:, 0] = -er[:, 0]
    return (np.concatenate([er[:, 0], el[:, 0]]), np.concatenate([er[:, 1] + 50, el[:, 1]]),
            np.concatenate([t[right], t[left]]))
