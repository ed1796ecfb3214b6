"""TEST INFRASTRUCTURE (CPU oracle; never imported by the product).
Restatement of the two scikit-image metrics render_test_views.py:34-35 calls, from their published definitions,
with scipy.ndimage.uniform_filter (the routine skimage.metrics.structural_similarity itself uses).
scikit-image is not installed in this image: parity with the library is unpinned."""
import numpy as np
from scipy.ndimage import uniform_filter


def peak_signal_noise_ratio(gt, pred, data_range=255.0):
    err = np.mean((gt.astype(np.float64) - pred.astype(np.float64)) ** 2)
    return 10 * np.log10(data_range ** 2 / err)


def structural_similarity(im1, im2, win_size=7, data_range=255.0, K1=0.01, K2=0.03):
    """multichannel=True, gaussian_weights=False, use_sample_covariance=True (the call's defaults)."""
    out = []
    for ch in range(im1.shape[-1]):
        x, y = im1[..., ch].astype(np.float64), im2[..., ch].astype(np.float64)
        NP = win_size ** 2
        cov_norm = NP / (NP - 1)
        ux, uy = uniform_filter(x, size=win_size), uniform_filter(y, size=win_size)
        uxx, uyy, uxy = uniform_filter(x * x, size=win_size), uniform_filter(y * y, size=win_size), uniform_filter(x * y, size=win_size)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        pad = (win_size - 1) // 2
        out.append(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return float(np.mean(out))
