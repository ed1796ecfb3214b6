"""Seeded synthetic inputs (cameras, default-init networks) used by bench.py, the tests and the golden
generator: the NeuMan dataset / checkpoints are not available offline."""
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
