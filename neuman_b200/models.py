"""Host-side mirror of the reference's network containers (models/vanilla.py:17-250,
models/human_nerf.py:20-90): same class names, constructor arguments, parameter names and shapes, so
reference checkpoints (`coarse_model_state_dict`, `hybrid_model_state_dict`, ...) load unchanged.
`forward()` runs the CUDA path; under autograd it runs the training kernel (activation stash) and returns
gradients to the network parameters (neuman_b200/autograd.py).
"""
import copy

import numpy as np
import torch
import torch.nn as nn

from . import ops


class Embedder(nn.Module):
    """models/vanilla.py:17-92. Holds only the description; the encoding itself is fused into the
    MLP kernels (csrc/nm_pe.cuh)."""

    def __init__(self, input_dims, max_freq, N_freqs, log_sampling=True, include_input=True, min_freq=0,
                 mapping='posenc'):
        super().__init__()
        if mapping not in ('posenc', 'rotate'):
            raise ValueError(mapping)
        if not log_sampling or not include_input:
            raise NotImplementedError("only log_sampling=True, include_input=True (the reference defaults)")
        self.input_dims = input_dims
        self.max_freq = max_freq
        self.min_freq = min_freq
        self.N_freqs = N_freqs
        self.log_sampling = log_sampling
        self.include_input = include_input
        self.mapping = mapping
        self.out_dim = input_dims + 2 * input_dims * N_freqs if mapping == 'posenc' else 3 + 6 * N_freqs

    def forward(self, inputs, cur_iter=None):
        raise NotImplementedError("Embedder is fused into the MLP kernel; call Joiner.forward")


class NeRF(nn.Module):
    """Parameter container with the reference's module names and shapes (models/vanilla.py:95-118):
    pts_linears.{0..depth-1}, views_linears.0, feature_linear, alpha_linear, rgb_linear (or output_linear).
    Modules are created in the reference's order so a seeded default init reproduces its weights."""

    def __init__(self, depth=8, width=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False,
                 scale=1.0, scale_type='no'):
        super().__init__()
        self.depth, self.width = depth, width
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips, self.use_viewdirs = skips, use_viewdirs
        self.scale, self.scale_type = scale, scale_type
        trunk = []
        fan_in = input_ch
        for layer in range(depth):
            trunk.append(nn.Linear(fan_in, width))
            # the layer after a skip index sees [encoded input, hidden] (input first, models/vanilla.py:131)
            fan_in = width + input_ch if layer in self.skips else width
        self.pts_linears = nn.ModuleList(trunk)
        if not use_viewdirs:
            self.output_linear = nn.Linear(width, output_ch)
            return
        self.views_linears = nn.ModuleList([nn.Linear(width + input_ch_views, width // 2)])
        self.feature_linear = nn.Linear(width, width)
        self.alpha_linear = nn.Linear(width, 1)
        self.rgb_linear = nn.Linear(width // 2, 3)

    def forward(self, input_pts, input_views=None):
        raise NotImplementedError("NeRF consumes encoded inputs; the fused CUDA path is Joiner.forward")


class Joiner(nn.Module):
    """models/vanilla.py:155-166."""

    def __init__(self, pos_pe, dir_pe, nerf):
        super().__init__()
        self.pos_pe, self.dir_pe, self.nerf = pos_pe, dir_pe, nerf

    def forward(self, input_pts, input_views=None):
        """input_pts [...,3], input_views [...,3] -> [...,4] = (r,g,b,sigma). CUDA only.
        Under autograd (a network parameter or an input requiring grad) the training kernel runs and the result
        carries gradients to the parameters and to input_pts / input_views (neuman_b200/autograd.py)."""
        if torch.is_grad_enabled() and input_views is not None and (
                any(p.requires_grad for p in self.nerf.parameters())
                or any(isinstance(t, torch.Tensor) and t.requires_grad for t in (input_pts, input_views))):
            from . import autograd
            return autograd.joiner_forward(self, input_pts, input_views)
        return ops.joiner_forward(self, input_pts, input_views)


def build_nerf(opt):
    """models/vanilla.py:208-250."""
    mapping = opt.posenc if hasattr(opt, 'posenc') else 'posenc'
    pos_pe = Embedder(opt.raw_pos_dim, opt.pos_max_freq, opt.pos_N_freqs, opt.log_sampling, opt.include_input,
                      min_freq=opt.pos_min_freq, mapping=mapping)
    dir_pe = Embedder(opt.raw_dir_dim, opt.dir_max_freq, opt.dir_N_freqs, opt.log_sampling, opt.include_input,
                      mapping=mapping)

    def one():
        return Joiner(pos_pe, dir_pe, NeRF(depth=opt.nerf_depth, width=opt.nerf_width, input_ch=pos_pe.out_dim,
                                           input_ch_views=dir_pe.out_dim, use_viewdirs=opt.use_viewdirs))
    coarse, fine = one(), one()
    if opt.use_cuda:
        coarse, fine = coarse.cuda(), fine.cuda()
    return coarse, fine


class HumanNeRF(nn.Module):
    """models/human_nerf.py:20-90: container of the background coarse/fine nets and the canonical
    human net (the offset nets are training-only, SURVEY.md §0.4, and not built here)."""

    def __init__(self, opt, poses=None, betas=None, alignments=None, scale=None):
        super().__init__()
        self.coarse_bkg_net, self.fine_bkg_net = build_nerf(opt)
        self.offset_nets = nn.ModuleList([])
        t = copy.deepcopy(opt)
        t.pos_min_freq = 0
        t.use_viewdirs = t.specular_can
        t.posenc = t.can_posenc
        self.coarse_human_net, _ = build_nerf(t)
        if poses is not None:
            self.poses = nn.Parameter(torch.from_numpy(np.asarray(poses)).float())
            self.betas = nn.Parameter(torch.from_numpy(np.asarray(betas)).float())
            self.alignments = nn.Parameter(torch.from_numpy(np.asarray(alignments)).float())
            self.scale = scale


def default_opt(**over):
    """options/options.py:52-81 defaults consumed by the path."""
    import types
    o = types.SimpleNamespace(
        use_cuda=torch.cuda.is_available(), nerf_depth=8, nerf_width=256, use_viewdirs=True, specular_can=True,
        raw_pos_dim=3, pos_min_freq=0, pos_max_freq=9, pos_N_freqs=10, raw_dir_dim=3, dir_max_freq=3, dir_N_freqs=4,
        log_sampling=True, include_input=True, can_posenc='rotate', rays_per_batch=2048, samples_per_ray=128,
        white_bkg=True, importance_samples_per_ray=128, num_offset_nets=0)
    o.__dict__.update(over)
    return o
