"""Drop-in: rebinds the reference's hot-path functions to the CUDA path.

    import neuman_b200; neuman_b200.install("/path/to/ml-neuman")
    # or: python -m neuman_b200.run render_test_views.py --scene_dir ...

After install(), `utils.render_utils.render_vanilla / render_smpl_nerf / render_hybrid_nerf /
render_hybrid_nerf_multi_persons / raw2outputs`, `utils.ray_utils.ray_to_samples /
ray_to_importance_samples / sample_pdf` and `models.vanilla.Joiner.forward` (inference, CUDA tensors,
grad disabled) run on libneuman_b200; the reference's nn.Modules, checkpoints and CLI scripts are
untouched.  Calls that need autograd (training) keep the reference's own torch implementation.
"""
import importlib
import sys

import torch


def install(reference_root=None, train=False):
    """train=True additionally routes the trainers' autograd path to the CUDA training kernels: Joiner.forward of
    8x256 nets (gradients to the parameters and to input_pts / input_views, which is what
    trainers/human_nerf_trainer.py:241-278 differentiates through) and raw2outputs with respect to `raw`
    (trainers/vanilla_nerf_trainer.py:45-96).  Everything else of the training graph (warp, offset nets, SMPL,
    regularisers) stays the reference's torch code."""
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from . import autograd, ops, render

    def on_cuda(*ts):
        return all(isinstance(t, torch.Tensor) and t.is_cuda for t in ts)

    def trainable_arch(j):
        n = j.nerf
        return (n.use_viewdirs and len(n.pts_linears) == 8 and tuple(n.skips) == (4,) and n.pts_linears[1].weight.shape == (256, 256))
    ru = importlib.import_module("utils.render_utils")
    ry = importlib.import_module("utils.ray_utils")
    mv = importlib.import_module("models.vanilla")

    def on_cuda_nograd(*ts):
        return (not torch.is_grad_enabled()) and all(isinstance(t, torch.Tensor) and t.is_cuda for t in ts)

    ref_forward = mv.Joiner.forward

    def joiner_forward(self, input_pts, input_views=None):
        if input_views is not None and on_cuda_nograd(input_pts, input_views) and self.nerf.use_viewdirs:
            return ops.joiner_forward(self, input_pts, input_views)
        if (train and input_views is not None and torch.is_grad_enabled() and on_cuda(input_pts, input_views)
                and trainable_arch(self)):
            return autograd.joiner_forward(self, input_pts, input_views)
        return ref_forward(self, input_pts, input_views)          # other training / CPU: reference torch path
    mv.Joiner.forward = joiner_forward

    ref_raw2outputs = ru.raw2outputs

    def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkg=True):
        if on_cuda_nograd(raw, z_vals, rays_d):
            return ops.raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkg)
        if (train and torch.is_grad_enabled() and on_cuda(raw, z_vals, rays_d)
                and not z_vals.requires_grad and not rays_d.requires_grad):
            return autograd.raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkg)
        return ref_raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkg)
    ru.raw2outputs = raw2outputs

    def cuda_model(m):
        return next(m.parameters()).is_cuda

    for name in ("render_vanilla", "render_smpl_nerf", "render_hybrid_nerf", "render_hybrid_nerf_multi_persons"):
        ref_fn, new_fn = getattr(ru, name), getattr(render, name)

        def make(ref_fn=ref_fn, new_fn=new_fn):
            def wrapped(model, *a, **k):
                if cuda_model(model):
                    return new_fn(model, *a, **k)
                return ref_fn(model, *a, **k)
            wrapped.__name__ = ref_fn.__name__
            return wrapped
        setattr(ru, name, make())

    ref_rts, ref_rtis, ref_pdf = ry.ray_to_samples, ry.ray_to_importance_samples, ry.sample_pdf

    def no_grad_inputs(*ts):
        return on_cuda(*ts) and (not torch.is_grad_enabled() or train)

    def ray_to_samples(ray_batch, samples_per_ray, lindisp=False, perturb=0., device='cpu', append_t=None):
        if append_t is None and no_grad_inputs(ray_batch['origin'], ray_batch['near']):
            return ops.ray_to_samples(ray_batch, samples_per_ray, lindisp, perturb)
        return ref_rts(ray_batch, samples_per_ray, lindisp, perturb, device, append_t)

    def ray_to_importance_samples(ray_batch, z_vals, weights, importance_samples_per_ray, device='cpu',
                                  including_old=True, append_t=None):
        if append_t is None and no_grad_inputs(z_vals, weights):       # samples are constants of the step (:150 detaches)
            return ops.ray_to_importance_samples(ray_batch, z_vals, weights, importance_samples_per_ray,
                                                 including_old=including_old)
        return ref_rtis(ray_batch, z_vals, weights, importance_samples_per_ray, device, including_old, append_t)

    def sample_pdf(bins, weights, N_samples, det=False, device='cpu'):
        if no_grad_inputs(bins, weights):
            return ops.sample_pdf(bins, weights, N_samples, det)
        return ref_pdf(bins, weights, N_samples, det, device)
    ry.ray_to_samples, ry.ray_to_importance_samples, ry.sample_pdf = ray_to_samples, ray_to_importance_samples, sample_pdf
    if train:
        # the human trainer's CPU libigl queries (utils/ray_utils.py:70, trainers/human_nerf_trainer.py:310,326)
        ref_diff = ry.warp_samples_to_canonical_diff

        def warp_samples_to_canonical_diff(pts, verts, faces, T):
            if isinstance(verts, torch.Tensor) and verts.is_cuda and isinstance(T, torch.Tensor) and T.is_cuda:
                T_inv, f_id, sd = ops.warp_samples_to_canonical_diff(pts, verts, faces, T)
                return T_inv, f_id.cpu().numpy(), sd.cpu().numpy()        # the reference returns igl's numpy arrays
            return ref_diff(pts, verts, faces, T)
        ry.warp_samples_to_canonical_diff = warp_samples_to_canonical_diff
        try:
            igl = importlib.import_module("igl")
            ref_sd = igl.signed_distance

            def signed_distance(P, V, F, *a, **k):
                if a or k or not torch.cuda.is_available():
                    return ref_sd(P, V, F, *a, **k)
                return ops.signed_distance(P, V, F)
            igl.signed_distance = signed_distance
        except ImportError:
            pass
    return {"render_utils": ru, "ray_utils": ry, "vanilla": mv}
