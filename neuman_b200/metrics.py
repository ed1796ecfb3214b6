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
decoder, the SSIM window statistics against cv2.blur's box filter.  LPIPS needs the AlexNet weights and is not built.
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


def eval_metrics(gts, preds):
    """render_test_views.py:27-41 without LPIPS: mean SSIM / PSNR over pairs of uint8 frames."""
    res = {'ssim': [], 'psnr': []}
    for gt, pred in zip(gts, preds):
        res['ssim'].append(ssim(pred, gt))
        res['psnr'].append(psnr(gt, pred))
    return {k: float(np.mean(v)) for k, v in res.items()}


def save_png(path, img):
    """imageio.imsave(save_path, out) (render_test_views.py:88): float or uint8 frame -> PNG via Pillow."""
    from PIL import Image
    Image.fromarray(to_uint8(img).cpu().numpy()).save(path)
