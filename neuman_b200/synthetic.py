"""Seeded synthetic inputs (cameras, default-init networks, an SMPL-shaped body model, the BASELINE.json configurations)
used by bench.py, the tests and the golden generators: the NeuMan dataset, its checkpoints and the licence-gated
SMPL_NEUTRAL.pkl are not available offline."""
import numpy as np
import torch


def camera(H, W, focal=None, seed=0, eye=(0.1, -0.05, -1.5), yaw=None):
    """Returns (K f64 3x3, c2w f32 4x4). Camera looks down +z (K^-1 [x,y,1] has z=1)."""
    rng = np.random.RandomState(seed)
    f = focal if focal is not None else 1000.0 * W / 1280
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]])
    a = rng.uniform(-0.2, 0.2) if yaw is None else yaw
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=np.float32)
    c2w[:3, 3] = np.asarray(eye, dtype=np.float32)
    return K, c2w


def seed_nets(build_nerf, opt, seed=1):
    """torch.manual_seed(seed); build_nerf(opt) -- the reference's and neuman_b200's builders create
    the same nn.Linear modules in the same order, so both get bit-identical default-init weights."""
    torch.manual_seed(seed)
    return build_nerf(opt)


def boost_density(joiner, gain=8.0, bias=0.3):
    """Default nn.Linear init leaves sigma ~ +-0.1 (and often <0 over a whole body); scale the alpha
    head so renders are not degenerate.  Applied identically to reference and product nets."""
    with torch.no_grad():
        joiner.nerf.alpha_linear.weight.mul_(gain)
        joiner.nerf.alpha_linear.bias.add_(bias)
    return joiner


def net_checksum(joiner):
    return float(sum(p.detach().double().abs().sum() for p in joiner.parameters()))


# ---------------------------------------------------------------------------------------------
# Deterministic synthetic SMPL-shaped body model.  The real SMPL_NEUTRAL.pkl is licence-gated and absent
# (README.md:60-71 of the reference).  A closed genus-0 lat-long mesh with exactly SMPL's counts -- 84 rings x 82
# segments + 2 poles = 6890 vertices, 13776 faces -- a 24-joint tree with SMPL's parent table, 4 non-zero skinning
# weights per vertex, 10 shape directions and a (zero) pose-blend basis, i.e. every key models/smpl.py:73-107 reads:
# f, v_template, shapedirs, J_regressor, posedirs, kintree_table, weights.
# ---------------------------------------------------------------------------------------------
N_RINGS, N_SEG = 84, 82
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21])

# rough SMPL rest joint locations (metres, y up)
_SMPL_JOINTS = np.array([
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
    d = np.linalg.norm(verts[:, None, :] - _SMPL_JOINTS[None], axis=2)      # [V,J]
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




# ---------------------------------------------------------------------------------------------
# BASELINE.json configurations 2-5 at their stated sizes: cameras, sample counts and actor placement.
# The posed bodies are produced from these descriptions either by the oracle (tests, goldens) or by the device SMPL
# kernels (bench.py); `window` is the 64x64 pixel block (x0, y0) whose 4096 rays the full-size parity tests compare
# with the reference's own output (tools/make_golden_fullsize.py picks it on a silhouette and stores it with the golden).
# ---------------------------------------------------------------------------------------------
FULLSIZE = {
    "cfg2": dict(driver="render_vanilla", H=720, W=1280, S=64, N=128, cam=dict(seed=1), near=0.0, far=3.14, actors=[]),
    "cfg3": dict(driver="render_smpl_nerf", H=512, W=512, S=128, N=0, near=0.0, far=1.0,
                 cam=dict(focal=1000.0 / 1280 * 512 * 3.0, seed=0, eye=(0.0, -0.05, -3.0), yaw=0.0),
                 actors=[dict(seed=1, center=(0.0, 0.0, 0.0), scale=0.4)]),
    "cfg4": dict(driver="render_hybrid_nerf", H=720, W=1280, S=128, N=128, cam=dict(seed=1), near=0.0, far=3.14,
                 actors=[dict(seed=1, center=(0.1, 0.0, 0.3), scale=0.45)]),
    "cfg5": dict(driver="render_hybrid_nerf_multi_persons", H=720, W=1280, S=128, N=128, cam=dict(seed=1), near=0.0, far=3.14,
                 actors=[dict(seed=1, center=(0.1, 0.0, 0.3), scale=0.45), dict(seed=4, center=(-0.35, 0.0, 0.6), scale=0.45),
                         dict(seed=7, center=(0.55, 0.05, 0.9), scale=0.45)]),
}


def fullsize_camera(name):
    c = FULLSIZE[name]
    return camera(c["H"], c["W"], **c["cam"])


def window_camera(K, x0, y0):
    """Intrinsics of the camera that sees only the pixel block starting at (x0, y0): the same rays as the full frame's
    pixels (x0 + x, y0 + y), so a reference renderer that only knows whole captures renders exactly that block."""
    Kw = np.array(K, dtype=np.float64).copy()
    Kw[0, 2] -= x0
    Kw[1, 2] -= y0
    return Kw


def actor_pose(desc):
    """Pose / shape / alignment of one synthetic actor (the random draws of oracle.synth_smpl.random_body)."""
    rng = np.random.RandomState(desc["seed"])
    pose = rng.normal(0, 0.3, size=(1, 72)).astype(np.float32)
    betas = rng.normal(0, 1.0, size=(1, 10)).astype(np.float32)
    align = np.eye(4)
    ang = rng.uniform(-0.3, 0.3)
    align[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    align = align.T.copy()                       # the reference applies alignment.T
    align[3, :3] = np.asarray(desc["center"]) / desc["scale"]
    return pose, betas, align
