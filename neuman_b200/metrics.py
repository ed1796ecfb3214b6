"""Frame output path of the evaluation script (render_test_views.py:27-41,85-92) without the
host round trips: float frame -> uint8, PSNR and SSIM as the script computes them through scikit-image, evaluated
with torch ops on whatever device the frames live on (float64 inside, like scikit-image).

    ssim(pred, gt, multichannel=True)              -> skimage.metrics.structural_similarity: 7x7 uniform window,
                                                      K1 = 0.01, K2 = 0.03, data_range 255 for uint8, sample
                                                      covariance (N/(N-1)), mean over the map cropped by 3 pixels
    skimage.metrics.peak_signal_noise_ratio(gt, pred) -> 10 log10(255^2 / mse)

scikit-image, imageio and lpips are not in this image: the formulas are restated from the published definitions
(oracle/metrics_oracle.py restates them once more with scipy's uniform_filter, the routine scikit-image itself
calls); parity with the libraries themselves is unpinned.  What IS pinned against third-party code present here
(tests/test_host.py::test_frame_metrics_against_opencv): PSNR against OpenCV's cv2.PSNR, the PNG file against OpenCV's
decoder, the SSIM window statistics against cv2.blur's box filter.  LPIPS: the network is restated below (class LPIPS);
its weights are not available offline.
"""
import numpy as np
import torch
import torch.nn.functional as F


def to_uint8(img):
    """float frame in [0,1] -> uint8 as imageio.imsave stores it (clip, scale by 255, round to nearest)."""
    t = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.asarray(img))
    if t.dtype == torch.uint8:
        return t
    return (t.double().clamp(0, 1) * 255.0 + 0.5).floor().to(torch.uint8)


def _as_f64(img, device=None):
    t = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.asarray(img))
    if device is not None:
        t = t.to(device)
    return t.double()


def psnr(gt, pred, data_range=255.0):
    """skimage.metrics.peak_signal_noise_ratio(gt, pred) for uint8 frames [H,W,3]."""
    a, b = _as_f64(gt), _as_f64(pred)
    mse = ((a - b.to(a.device)) ** 2).mean()
    return float(10.0 * torch.log10(data_range ** 2 / mse))


def ssim(pred, gt, data_range=255.0, win_size=7, K1=0.01, K2=0.03):
    """skimage.metrics.structural_similarity(pred, gt, multichannel=True) for uint8 frames [H,W,C]: mean over the
    channels of the mean SSIM of each channel."""
    x, y = _as_f64(pred), _as_f64(gt)
    y = y.to(x.device)
    x, y = x.permute(2, 0, 1)[None], y.permute(2, 0, 1)[None]              # [1,C,H,W]
    NP = win_size * win_size
    cov_norm = NP / (NP - 1.0)                                            # use_sample_covariance=True

    def box(t):                                                           # uniform_filter, then the crop by (win-1)//2:
        return F.avg_pool2d(t, win_size, stride=1)                        # exactly the 'valid' window means
    ux, uy = box(x), box(y)
    vx = cov_norm * (box(x * x) - ux * ux)
    vy = cov_norm * (box(y * y) - uy * uy)
    vxy = cov_norm * (box(x * y) - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return float(S.mean(dim=(2, 3)).mean())


def eval_metrics(gts, preds, lpips_fn=None):
    """render_test_views.py:27-41: mean SSIM / PSNR (and LPIPS when an `LPIPS` module with loaded weights is given, :36-38)
    over pairs of uint8 frames."""
    res = {'ssim': [], 'psnr': []}
    if lpips_fn is not None:
        res['lpips'] = []
    for gt, pred in zip(gts, preds):
        res['ssim'].append(ssim(pred, gt))
        res['psnr'].append(psnr(gt, pred))
        if lpips_fn is not None:
            dev = next(lpips_fn.parameters()).device
            a, b = (_as_f64(x, dev).float().permute(2, 0, 1)[None] / 127.5 - 1 for x in (pred, gt))     # np_img_to_torch_img(.)/127.5-1
            with torch.no_grad():
                res['lpips'].append(float(lpips_fn(a, b)[0, 0, 0, 0]))
    return {k: float(np.mean(v)) for k, v in res.items()}


# ---------------------------------------------------------------------------------------------
# LPIPS (AlexNet variant): the perceptual term of render_test_views.py:19,36-38 and of the human trainer's patch loss
# (trainers/human_nerf_trainer.py:152-153,432-435: `lpips.LPIPS(net='alex')`).
#
# The `lpips` package and its weights are not in this image and cannot be fetched.  This class restates the network from
# its published definition (Zhang et al., "The Unreasonable Effectiveness of Deep Features as a Perceptual Metric", 2018;
# lpips v0.1): input shift / scale, the five ReLU outputs of torchvision's AlexNet `features`, unit-normalisation over
# channels, squared difference, a learned non-negative 1x1 convolution per layer, spatial mean, sum over layers.  Module and
# parameter names are the package's (`scaling_layer.shift/scale`, `net.slice{1..5}.{0,3,6,8,10}.weight/bias`,
# `lin{0..4}.model.1.weight`, mirrored under `lins.{k}`), so `load_state_dict(lpips.LPIPS(net='alex').state_dict())` works
# where the package exists.  It is evaluated with library convolutions (cuDNN) -- a 32x32 patch or one frame per call, not
# a kernel of this repo; parity with the package is UNPINNED (no weights here): tests/test_host.py checks the feature stack
# against torchvision's AlexNet with the same random weights and the distance against a numpy restatement.
# ---------------------------------------------------------------------------------------------
class _AlexFeatures(torch.nn.Module):
    """torchvision.models.alexnet().features cut at its five ReLUs (lpips/pretrained_networks.py: slices [0:2], [2:5], [5:8],
    [8:10], [10:12], sub-modules named by their index in `features`)."""

    def __init__(self):
        super().__init__()
        nn = torch.nn
        def seq(items):
            s = nn.Sequential()
            for name, mod in items:
                s.add_module(str(name), mod)
            return s
        self.slice1 = seq([(0, nn.Conv2d(3, 64, 11, 4, 2)), (1, nn.ReLU())])
        self.slice2 = seq([(2, nn.MaxPool2d(3, 2)), (3, nn.Conv2d(64, 192, 5, 1, 2)), (4, nn.ReLU())])
        self.slice3 = seq([(5, nn.MaxPool2d(3, 2)), (6, nn.Conv2d(192, 384, 3, 1, 1)), (7, nn.ReLU())])
        self.slice4 = seq([(8, nn.Conv2d(384, 256, 3, 1, 1)), (9, nn.ReLU())])
        self.slice5 = seq([(10, nn.Conv2d(256, 256, 3, 1, 1)), (11, nn.ReLU())])

    def forward(self, x):
        outs = []
        for s in (self.slice1, self.slice2, self.slice3, self.slice4, self.slice5):
            x = s(x)
            outs.append(x)
        return outs


class _NetLinLayer(torch.nn.Module):
    def __init__(self, chn_in):
        super().__init__()
        self.model = torch.nn.Sequential(torch.nn.Dropout(), torch.nn.Conv2d(chn_in, 1, 1, bias=False))

    def forward(self, x):
        return self.model(x)


class _ScalingLayer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('shift', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, x):
        return (x - self.shift) / self.scale


class LPIPS(torch.nn.Module):
    """lpips.LPIPS(net='alex') (v0.1, spatial=False, lpips=True).  forward(in0, in1, normalize=False): images [N,3,H,W] in
    [-1, 1] (in [0, 1] with normalize=True) -> [N,1,1,1].  Weights: load_state_dict of the package's module, or
    `load_pretrained(alexnet_features_state_dict, lin_state_dict)` from torchvision's alexnet checkpoint (`features.*`) and the
    package's `weights/v0.1/alex.pth` (`lin{k}.model.1.weight`).  Evaluation mode by default (the package's default too)."""
    CHNS = (64, 192, 384, 256, 256)

    def __init__(self):
        super().__init__()
        self.scaling_layer = _ScalingLayer()
        self.net = _AlexFeatures()
        for k, c in enumerate(self.CHNS):
            setattr(self, f'lin{k}', _NetLinLayer(c))
        self.lins = torch.nn.ModuleList([getattr(self, f'lin{k}') for k in range(5)])
        self.pretrained = False
        self.eval()

    def load_pretrained(self, alexnet_state, lin_state):
        feats = {k[len('features.'):]: v for k, v in alexnet_state.items() if k.startswith('features.')}
        own = self.net.state_dict()
        for k in own:                                        # 'slice2.3.weight' <- 'features.3.weight'
            own[k] = feats[k.split('.', 1)[1]]
        self.net.load_state_dict(own)
        for k in range(5):
            getattr(self, f'lin{k}').model[1].weight.data.copy_(lin_state[f'lin{k}.model.1.weight'])
        self.pretrained = True
        return self

    def load_state_dict(self, state_dict, strict=True):
        out = super().load_state_dict(state_dict, strict=strict)
        self.pretrained = True
        return out

    @staticmethod
    def _unit(x, eps=1e-10):
        return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)

    def forward(self, in0, in1, normalize=False):
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        f0, f1 = self.net(self.scaling_layer(in0)), self.net(self.scaling_layer(in1))
        val = 0
        for k in range(5):
            d = (self._unit(f0[k]) - self._unit(f1[k])) ** 2
            val = val + self.lins[k](d).mean(dim=(2, 3), keepdim=True)
        return val


def lpips_patch_loss(lpips_fn, rgb_map, color, patch_size=32):
    """The LPIPS term of the human trainer (trainers/human_nerf_trainer.py:432-435): the first patch_size^2 rays of the
    batch are a patch in row-major order (datasets/human_rays.py:163-183, neuman_b200.data.HumanRayBatcher); both in [0,1]."""
    n = patch_size * patch_size
    a = rgb_map[:n].reshape(patch_size, patch_size, -1).permute(2, 0, 1) * 2 - 1
    b = color[:n].to(rgb_map.device).reshape(patch_size, patch_size, -1).permute(2, 0, 1) * 2 - 1
    out = lpips_fn(a, b)
    assert out.numel() == 1
    return out.flatten()[0]


def save_png(path, img):
    """imageio.imsave(save_path, out) (render_test_views.py:88): float or uint8 frame -> PNG via Pillow."""
    from PIL import Image
    Image.fromarray(to_uint8(img).cpu().numpy()).save(path)
