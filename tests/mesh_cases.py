"""Closest-point cases whose answers are known BY CONSTRUCTION, independent of any closest-point code (test data).

libigl (the reference's closest-point library, utils/ray_utils.py:53,55,70; environment.yml:13) is not in the reference
tree nor in this image, so the oracle restatement (oracle/mesh_oracle.py) cannot be pinned against it.  These cases pin
the *semantics* instead -- exact Euclidean closest point on a triangle soup -- from first principles:

  * `constructed_cases`: pick a feature point q on a mesh (face interior, edge, vertex) and move away from it inside the
    normal cone of that feature: for a CONVEX closed mesh every point of q + cone(q) has q as its unique closest point, so
    (sqrD, C, barycentrics, sign) are known without running any search;
  * `icosphere`: |signed distance| of a unit icosphere differs from | |p| - 1 | by at most the sagitta of its faces;
  * `shared_edge_cases`: the closest point lies on an edge shared by two faces -- whichever of the two a library reports,
    the blended transform of the warp is the same (only the two edge vertices carry weight).
"""
import numpy as np


def cube():
    """Unit cube [-1,1]^3, 12 outward-wound triangles."""
    V = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    F = []
    for a, b, c, d in quads:
        F += [(a, b, c), (a, c, d)]
    F = np.array(F, dtype=np.int64)
    # make every face outward (centroid . normal > 0 for a body centred at the origin)
    for i, (a, b, c) in enumerate(F):
        n = np.cross(V[b] - V[a], V[c] - V[a])
        if np.dot(n, V[a] + V[b] + V[c]) < 0:
            F[i] = (a, c, b)
    return V, F


def icosphere(level=2):
    t = (1 + 5 ** 0.5) / 2
    V = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    V = [np.array(v, dtype=np.float64) / np.linalg.norm(v) for v in V]
    for _ in range(level):
        cache, F2 = {}, []

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = V[i] + V[j]
                V.append(m / np.linalg.norm(m))
                cache[key] = len(V) - 1
            return cache[key]
        for a, b, c in F:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            F2 += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        F = F2
    return np.array(V), np.array(F, dtype=np.int64)


def _face_normals(V, F):
    n = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    return n / np.linalg.norm(n, axis=1, keepdims=True)


def constructed_cases(V, F, seed=0, per_kind=64, tmax=0.7):
    """For a convex, outward-wound closed mesh.  Returns a dict of arrays:
    P [n,3] query points (outside), C [n,3] their closest points, D [n] distances, kind [n] (0 face, 1 edge, 2 vertex),
    faces_ok: list of sets of face indices that contain C (any of them is a correct answer for I)."""
    rng = np.random.RandomState(seed)
    fn = _face_normals(V, F)
    edges = {}
    for f, (a, b, c) in enumerate(F):
        for e in ((a, b), (b, c), (c, a)):
            edges.setdefault((min(e), max(e)), []).append(f)
    vfaces = {}
    for f, tri in enumerate(F):
        for v in tri:
            vfaces.setdefault(int(v), []).append(f)
    P, C, D, kind, ok = [], [], [], [], []
    for _ in range(per_kind):                                  # face interiors: along the face normal
        f = rng.randint(len(F))
        w = rng.dirichlet((1.5, 1.5, 1.5))
        q = w @ V[F[f]]
        t = rng.uniform(0.01, tmax)
        P.append(q + t * fn[f]); C.append(q); D.append(t); kind.append(0); ok.append({f})
    ekeys = sorted(edges)
    for _ in range(per_kind):                                  # edges: between the two adjacent face normals
        i, j = ekeys[rng.randint(len(ekeys))]
        f0, f1 = edges[(i, j)]
        s = rng.uniform(0.05, 0.95)
        q = (1 - s) * V[i] + s * V[j]
        a = rng.uniform(0.05, 0.95)
        d = (1 - a) * fn[f0] + a * fn[f1]
        d /= np.linalg.norm(d)
        t = rng.uniform(0.01, tmax)
        P.append(q + t * d); C.append(q); D.append(t); kind.append(1); ok.append({f0, f1})
    for _ in range(per_kind):                                  # vertices: inside the cone of the adjacent face normals
        v = rng.randint(len(V))
        w = rng.dirichlet(np.ones(len(vfaces[v])))
        d = w @ fn[vfaces[v]]
        d /= np.linalg.norm(d)
        t = rng.uniform(0.01, tmax)
        P.append(V[v] + t * d); C.append(V[v].copy()); D.append(t); kind.append(2); ok.append(set(vfaces[v]))
    return {"P": np.array(P), "C": np.array(C), "D": np.array(D), "kind": np.array(kind), "faces_ok": ok}


def inside_cases(V, F, seed=0, n=64):
    """Points INSIDE the cube [-1,1]^3: distance = 1 - max|coordinate|, sign negative, closest point = the projection
    onto the nearest side (unique when the largest |coordinate| is unique)."""
    rng = np.random.RandomState(seed)
    P = rng.uniform(-0.9, 0.9, (n, 3))
    ax = np.abs(P).argmax(1)
    srt = np.sort(np.abs(P), 1)
    keep = (srt[:, 2] - srt[:, 1]) > 0.05                      # stay away from the medial axis
    P, ax = P[keep], ax[keep]
    C = P.copy()
    C[np.arange(len(P)), ax] = np.sign(P[np.arange(len(P)), ax])
    return {"P": P, "C": C, "D": 1 - np.abs(P).max(1)}


def single_triangle_cases(seed=0, n=200):
    """One (non-convex-irrelevant) triangle in general position: region answers from the in-plane decomposition.
    q = projection of p onto the plane; inside -> q; else the closest point of the triangle's boundary, found by clamping
    the projection onto each edge LINE to the segment and taking the nearest of the three (no Voronoi-region logic)."""
    rng = np.random.RandomState(seed)
    A, B, Cc = rng.normal(0, 1, (3, 3))
    P = rng.normal(0, 1.5, (n, 3))
    nrm = np.cross(B - A, Cc - A)
    nrm /= np.linalg.norm(nrm)
    q = P - ((P - A) @ nrm)[:, None] * nrm
    M = np.stack([B - A, Cc - A], 1)                           # 3x2: solve the in-plane coordinates
    st = np.linalg.lstsq(M, (q - A).T, rcond=None)[0].T
    inside = (st[:, 0] >= 0) & (st[:, 1] >= 0) & (st.sum(1) <= 1)
    best = np.full(n, np.inf)
    Cl = q.copy()
    for u, v in ((A, B), (B, Cc), (Cc, A)):
        t = np.clip(((P - u) @ (v - u)) / ((v - u) @ (v - u)), 0, 1)
        c = u + t[:, None] * (v - u)
        d = np.linalg.norm(P - c, axis=1)
        upd = d < best
        best = np.where(upd, d, best)
        Cl = np.where((upd & ~inside)[:, None], c, Cl)
    D = np.where(inside, np.abs((P - A) @ nrm), best)
    return {"V": np.stack([A, B, Cc]), "F": np.array([[0, 1, 2]]), "P": P, "C": Cl, "D": D}


def shared_edge_cases(seed=0, n=64):
    """Two triangles sharing the edge (v0, v1), folded like a roof; query points above the ridge inside the wedge of the
    two face normals: the closest point is on the shared edge and BOTH faces are exact arg-mins."""
    rng = np.random.RandomState(seed)
    V = np.array([[0, 0, 0], [1, 0, 0], [0.4, 1, -0.6], [0.6, -1, -0.6]], dtype=np.float64)
    F = np.array([[0, 1, 2], [1, 0, 3]], dtype=np.int64)
    fn = _face_normals(V, F)
    s = rng.uniform(0.1, 0.9, n)
    a = rng.uniform(0.1, 0.9, n)
    d = (1 - a)[:, None] * fn[0] + a[:, None] * fn[1]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    t = rng.uniform(0.05, 0.5, n)
    q = (1 - s)[:, None] * V[0] + s[:, None] * V[1]
    return {"V": V, "F": F, "P": q + t[:, None] * d, "C": q, "D": t, "s": s}
