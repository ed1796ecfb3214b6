"""TEST INFRASTRUCTURE ONLY -- default option namespace of the reference (options/options.py:52-81,
train.py:223-224, render_360.py:144-151) used to build reference nets for oracle validation."""
import types


def default_opt(**over):
    o = types.SimpleNamespace(
        use_cuda=False, nerf_depth=8, nerf_width=256, use_viewdirs=True, specular_can=True,
        raw_pos_dim=3, pos_min_freq=0, pos_max_freq=9, pos_N_freqs=10, raw_dir_dim=3,
        dir_max_freq=3, dir_N_freqs=4, log_sampling=True, include_input=True, can_posenc='rotate',
        rays_per_batch=2048, samples_per_ray=128, white_bkg=True, importance_samples_per_ray=128,
        num_offset_nets=1, offset_scale=1.0, offset_scale_type='linear',
        out_dir='/nonexistent', load_background='none', load_can='none')
    o.__dict__.update(over)
    return o
