"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic SMPL-shaped body model.

The real SMPL_NEUTRAL.pkl is licence-gated and absent (README.md:60-71 of the reference). This
builds a closed genus-0 lat-long mesh with exactly SMPL's counts -- 84 rings x 82 segments + 2 poles
= 6890 vertices, 13776 faces -- a 24-joint tree with SMPL's parent table, 4 non-zero skinning weights
per vertex, 10 shape directions and a (zero) pose-blend basis, i.e. every key models/smpl.py:73-107
reads: f, v_template, shapedirs, J_regressor, posedirs, kintree_table, weights.
"""
import pickle

import numpy as np

from neuman_b200.synthetic import N_RINGS, N_SEG, SMPL_PARENTS, make_model      # noqa: F401  (the synthetic model itself is shared with bench.py)


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
