"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic SMPL-shaped body model.

The real SMPL_NEUTRAL.pkl is licence-gated and absent (README.md:60-71 of the reference). This
builds a closed genus-0 lat-long mesh with exactly SMPL's counts -- 84 rings x 82 segments + 2 poles
= 6890 vertices, 13776 faces -- a 24-joint tree with SMPL's parent table, 4 non-zero skinning weights
per vertex, 10 shape directions and a (zero) pose-blend basis, i.e. every key models/smpl.py:73-107
reads: f, v_template, shapedirs, J_regressor, posedirs, kintree_table, weights.
"""
import pickle

import numpy as np

N_RINGS, N_SEG = 84, 82
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21])

# rough SMPL rest joint locations (metres, y up)
_JOINTS = np.array([
    [0.00, 0.00, 0.00], [0.07, -0.09, 0.00], [-0.07, -0.09, 0.00], [0.00, 0.11, -0.02],
    [0.10, -0.47, 0.00], [-0.10, -0.47, 0.00], [0.00, 0.25, 0.00], [0.09, -0.87, -0.03],
    [-0.09, -0.87, -0.03], [0.00, 0.30, 0.02], [0.11, -0.93, 0.09], [-0.11, -0.93, 0.09],
    [0.00, 0.51, -0.02], [0.08, 0.42, -0.01], [-0.08, 0.42, -0.01], [0.00, 0.60, 0.03],
    [0.17, 0.44, -0.02], [-0.17, 0.44, -0.02], [0.43, 0.43, -0.03], [-0.43, 0.43, -0.03],
    [0.68, 0.43, -0.03], [-0.68, 0.43, -0.03], [0.76, 0.42, -0.04], [-0.76, 0.42, -0.04]])


def _mesh():
    ys = np.linspace(-1.0, 0.72, N_RINGS + 2)[1:-1]                 # ring heights
    prof = 0.08 + 0.10 * np.exp(-((ys - 0.15) / 0.35) ** 2) + 0.05 * np.exp(-((ys + 0.55) / 0.3) ** 2)
    ang = np.linspace(0, 2 * np.pi, N_SEG, endpoint=False)
    verts = [[0.0, -1.0, 0.0]]
    for y, r in zip(ys, prof):
        for a in ang:
            verts.append([r * np.cos(a) * 1.25, y, r * np.sin(a) * 0.8])
    verts.append([0.0, 0.72, 0.0])
    verts = np.asarray(verts, dtype=np.float64)
    faces = []
    top = len(verts) - 1

    def vid(ring, seg):
        return 1 + ring * N_SEG + (seg % N_SEG)
    for s in range(N_SEG):
        faces.append([0, vid(0, s + 1), vid(0, s)])
        faces.append([top, vid(N_RINGS - 1, s), vid(N_RINGS - 1, s + 1)])
    for r in range(N_RINGS - 1):
        for s in range(N_SEG):
            a, b, c, d = vid(r, s), vid(r, s + 1), vid(r + 1, s), vid(r + 1, s + 1)
            faces.append([a, b, c])
            faces.append([b, d, c])
    faces = np.asarray(faces, dtype=np.int64)
    assert verts.shape == (6890, 3) and faces.shape == (13776, 3)
    return verts, faces


def make_model(seed=0):
    """Returns a dict of numpy arrays with the SMPL pickle keys."""
    rng = np.random.RandomState(seed)
    verts, faces = _mesh()
    nj = 24
    d = np.linalg.norm(verts[:, None, :] - _JOINTS[None], axis=2)      # [V,J]
    near4 = np.argsort(d, axis=1)[:, :4]
    w = np.zeros((verts.shape[0], nj))
    rows = np.arange(verts.shape[0])[:, None]
    w[rows, near4] = 1.0 / (d[rows, near4] + 0.05) ** 2
    w /= w.sum(1, keepdims=True)
    # joint regressor: softmax of negative distance over the 32 nearest vertices of each joint
    jr = np.zeros((nj, verts.shape[0]))
    for j in range(nj):
        idx = np.argsort(d[:, j])[:32]
        ww = np.exp(-d[idx, j] * 20.0)
        jr[j, idx] = ww / ww.sum()
    # smooth low-frequency shape directions
    sd = np.zeros((verts.shape[0], 3, 10))
    for k in range(10):
        f = rng.uniform(1.0, 4.0, size=3)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        amp = rng.uniform(0.003, 0.012)
        sd[:, :, k] = amp * np.sin(verts * f[None] + ph[None]) * (verts / (np.abs(verts).max(0) + 1e-9))
    kin = np.stack([SMPL_PARENTS.copy(), np.arange(nj)]).astype(np.int64)
    kin[0, 0] = 2 ** 32 - 1
    return {
        "f": faces.astype(np.uint32),
        "v_template": verts,
        "shapedirs": sd,
        "J_regressor": jr,
        "posedirs": np.zeros((verts.shape[0], 3, 207)),
        "kintree_table": kin,
        "weights": w,
    }


def write_pickle(path, seed=0):
    with open(path, "wb") as fp:
        pickle.dump(make_model(seed), fp, protocol=2)


def torch_model(seed=0):
    """The dict oracle.neuman_oracle.smpl_lbs consumes (float32 tensors, parents[0] = -1)."""
    import torch
    m = make_model(seed)
    parents = torch.from_numpy(m["kintree_table"][0].astype(np.int64)).clone()
    parents[0] = -1
    return {
        "v_template": torch.from_numpy(m["v_template"]).float(),
        "shapedirs": torch.from_numpy(m["shapedirs"]).float(),
        "J_regressor": torch.from_numpy(m["J_regressor"]).float(),
        "parents": parents,
        "weights": torch.from_numpy(m["weights"]).float(),
        "faces": m["f"].astype(np.int64),
    }


def random_body(seed=0, scale=0.4, center=(0.0, 0.0, 0.0)):
    """A posed body in 'scene' units: returns dict(verts f32 [6890,3], faces int64 [13776,3],
    Ts f64 [6914,4,4], joints f32 [24,3], da_verts, geo_threshold). Mirrors SURVEY §8(d):
    pose ~ N(0,0.3^2), beta ~ N(0,1), scale so the body spans ~`scale`*1.7 scene units."""
    import torch
    from oracle import neuman_oracle as no
    rng = np.random.RandomState(seed)
    model = torch_model(0)
    pose = torch.from_numpy(rng.normal(0, 0.3, size=(1, 72))).float()
    betas = torch.from_numpy(rng.normal(0, 1.0, size=(1, 10))).float()
    align = np.eye(4)
    ang = rng.uniform(-0.3, 0.3)
    align[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    align = align.T.copy()                       # the reference applies alignment.T
    align[3, :3] = np.asarray(center) / scale    # so that alignment.T carries the translation
    verts, joints, Ts, da_verts = no.scene_vertex_transforms(model, pose, betas, align, scale)
    geo = float(np.linalg.norm(joints[3] - joints[0]))   # render_test_views.py:60-64 (one frame)
    return {"verts": verts, "faces": model["faces"], "Ts": Ts, "joints": joints,
            "da_verts": da_verts, "geo_threshold": geo, "pose": pose.numpy(), "betas": betas.numpy()}
