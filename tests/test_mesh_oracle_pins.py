"""Pins of the closest-point oracle (oracle/mesh_oracle.py) that do not depend on any closest-point code: cases whose
answers are known by construction (tests/mesh_cases.py), a numerical optimiser (scipy SLSQP over the barycentric simplex of
every triangle) and the analytic sphere.  libigl itself is absent (parity with it stays unpinned); these
tests pin the semantics the restatement claims -- exact Euclidean closest point, barycentrics in vertex order, negative
inside -- from first principles."""
import numpy as np

from oracle import mesh_oracle as mo
from tests import mesh_cases as mc


def _check(case, V, F, sqrD, I, C, tol=1e-12):
    assert np.abs(C - case["C"]).max() < tol
    assert np.abs(sqrD - case["D"] ** 2).max() < tol
    if "faces_ok" in case:
        assert all(int(i) in ok for i, ok in zip(I, case["faces_ok"]))


def test_constructed_cases_on_convex_meshes():
    for V, F in (mc.cube(), mc.icosphere(1)):
        case = mc.constructed_cases(V, F, seed=1, per_kind=96)
        sqrD, I, C = mo.point_mesh_squared_distance(case["P"], V, F)
        _check(case, V, F, sqrD, I, C)
        # barycentrics of the closest point reproduce it, sum to one, and vanish off the feature
        L = mo.barycentric_coordinates_tri(C, V[F[I, 0]], V[F[I, 1]], V[F[I, 2]])
        assert np.abs(L.sum(1) - 1).max() < 1e-12
        assert np.abs(np.einsum("ik,ikj->ij", L, V[F[I]]) - C).max() < 1e-12
        assert (L > -1e-12).all()
        k = case["kind"]
        assert ((L > 1e-9).sum(1)[k == 2] == 1).all() and ((L > 1e-9).sum(1)[k == 1] == 2).all()
        # outside a closed outward-wound mesh: positive sign, on faces, edges and vertices alike
        S, I2, C2 = mo.signed_distance(case["P"], V, F)
        assert (S > 0).all() and np.abs(S - case["D"]).max() < 1e-12


def test_inside_the_cube_is_negative():
    V, F = mc.cube()
    case = mc.inside_cases(V, F, seed=2, n=200)
    S, I, C = mo.signed_distance(case["P"], V, F)
    assert (S < 0).all()
    assert np.abs(-S - case["D"]).max() < 1e-12 and np.abs(C - case["C"]).max() < 1e-12


def test_single_triangle_all_regions():
    case = mc.single_triangle_cases(seed=3, n=400)
    sqrD, I, C = mo.point_mesh_squared_distance(case["P"], case["V"], case["F"])
    assert np.abs(np.sqrt(sqrD) - case["D"]).max() < 1e-12
    assert np.abs(C - case["C"]).max() < 1e-10


def test_against_a_numerical_optimiser():
    """min over faces of min_{l in simplex} |l.V_f - p|^2 by SLSQP, on a random (non-convex) soup."""
    from scipy.optimize import minimize
    rng = np.random.RandomState(5)
    V = rng.normal(0, 1, (18, 3))
    F = np.array([rng.choice(18, 3, replace=False) for _ in range(14)])
    P = rng.normal(0, 1.5, (40, 3))
    sqrD, I, C = mo.point_mesh_squared_distance(P, V, F)
    cons = ({"type": "eq", "fun": lambda l: l.sum() - 1.0},)
    for p, d2, i in zip(P, sqrD, I):
        best = np.inf
        for f in F:
            tri = V[f]
            for x0 in (np.ones(3) / 3, np.array([.8, .1, .1]), np.array([.1, .8, .1]), np.array([.1, .1, .8])):
                r = minimize(lambda l: ((l @ tri - p) ** 2).sum(), x0, jac=lambda l: 2 * tri @ (l @ tri - p),
                             bounds=[(0, 1)] * 3, constraints=cons, method="SLSQP", options={"ftol": 1e-15, "maxiter": 200})
                best = min(best, r.fun)
        assert abs(best - d2) < 1e-8 * (1 + d2), (best, d2)
        assert d2 <= best * (1 + 1e-10) + 1e-12                     # never worse than the optimiser (SLSQP meets its constraints to ~1e-13)


def test_sphere_signed_distance():
    V, F = mc.icosphere(3)                                          # 1280 faces, edge ~0.16, sagitta ~3.2e-3
    rng = np.random.RandomState(7)
    P = rng.normal(0, 1, (500, 3))
    P *= (rng.uniform(0.3, 1.8, 500) / np.linalg.norm(P, axis=1))[:, None]
    S, I, C = mo.signed_distance(P, V, F)
    r = np.linalg.norm(P, axis=1)
    sag = 1 - np.linalg.norm(V[F].mean(1), axis=1).min()
    assert sag < 5e-3
    clear = np.abs(r - 1) > 2 * sag
    assert (np.sign(S[clear]) == np.sign(r[clear] - 1)).all()
    # the mesh is inscribed in the unit sphere: its surface lies between radius 1 - sag and 1
    assert (S >= r - 1 - 1e-12).all() and (S <= r - 1 + sag * 1.0001 + 1e-12).all()


def test_shared_edge_ties_do_not_change_the_warp():
    """Both faces of a shared edge are exact arg-mins; whichever one a library reports, the blended transform
    (utils/ray_utils.py:56) is the same because only the two edge vertices carry weight."""
    case = mc.shared_edge_cases(seed=9)
    V, F, P = case["V"], case["F"], case["P"]
    sqrD, I, C = mo.point_mesh_squared_distance(P, V, F)
    assert np.abs(C - case["C"]).max() < 1e-12 and (I == 0).all()   # this restatement: lowest face index
    T = np.random.RandomState(1).normal(0, 1, (4, 4, 4))
    blends = []
    for f in (0, 1):
        L = mo.barycentric_coordinates_tri(C, *(np.repeat(V[F[f, k]][None], len(P), 0) for k in range(3)))
        blends.append(np.einsum("nk,kij->nij", L, T[F[f]]))
    assert np.abs(blends[0] - blends[1]).max() < 1e-12
