"""Generates tests/golden/batches.npz: ray batches produced by the UNMODIFIED reference dataset classes
(datasets/background_rays.py BackgroundRayDataset.__getitem__, datasets/human_rays.py HumanRayDataset.__getitem__, with
and without the 32x32 LPIPS patch branch) on seeded synthetic captures.  Run in the build container only:

    python tools/make_golden_batches.py

The reference draws its pixels with `random` / `np.random` inside __getitem__ and does not return them; this script wraps
`utils.ray_utils.shot_rays` (the one function every sampled pixel list passes through) with a recorder, so the committed
fixture holds (captures, pixels, batch): tests/test_gpu_train.py feeds the same pixels to neuman_b200.data and compares the
batches.  The scene object is a stand-in with the four members the datasets read (captures, __getitem__,
fname_to_index_dict, get_captures_by_view_id); the reference's scene reader (data_io/neuman_helper.py) parses COLMAP /
dataset folders and is outside the path.
"""
import contextlib
import io
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, scenes, synth_smpl      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "batches.npz")


def make_caps(ref, n=3, seed=3, human=False):
    rng = np.random.RandomState(seed)
    caps = []
    for k in range(n):
        H, W = 48 + 8 * k, 64
        K, c2w = scenes.camera(H, W, seed=10 + k)
        cam = ref.pinhole_camera.PinholeCamera(W, H, K[0, 0], K[1, 1], K[0, 2], K[1, 2])
        pose = ref.camera_pose.CameraPose.from_camera_to_world(c2w.astype(np.float64))
        cap = ref.captures.BasePinholeCapture(cam, pose)
        cap.near, cap.far = {"bkg": 0.3 + 0.1 * k, "human": 0.5}, {"bkg": 4.0 + k, "human": 5.0}
        cap.image = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
        cap.image_path = f"/synthetic/{k:05d}.png"
        cap.depth_map = rng.uniform(0.5, 3.0, (H, W)).astype(np.float32)
        cap.fused_depth_map = rng.uniform(0.5, 3.0, (H, W)).astype(np.float32)
        mask = np.zeros((H, W), np.uint8)
        mask[6 + k:40, 10 + 2 * k:50] = 1
        mask[rng.uniform(size=(H, W)) < 0.05] = 1
        cap.mask = mask
        cap.binary_mask = mask.copy()
        if human or k == 1:
            border = np.zeros((H, W), np.uint8)
            border[4 + k:42, 8 + 2 * k:52] = 1
            border[mask != 0] = 0
            cap.border_mask = border
        cap.frame_id = {"frame_id": k, "total_frames": 7}
        cap.posed_mesh = types.SimpleNamespace(device="cpu")
        caps.append(cap)
    return caps


class FakeScene:
    def __init__(self, caps):
        self.captures = caps
        self.fname_to_index_dict = {os.path.basename(c.image_path): i for i, c in enumerate(caps)}

    def __getitem__(self, fname):
        return self.captures[self.fname_to_index_dict[fname]]

    def get_captures_by_view_id(self, i):
        return [self.captures[i]]


def cap_arrays(caps, prefix):
    g = {}
    for i, c in enumerate(caps):
        p = f"{prefix}{i}_"
        g[p + "K"], g[p + "c2w"] = np.asarray(c.intrinsic_matrix), np.asarray(c.cam_pose.camera_to_world)
        g[p + "image"], g[p + "mask"], g[p + "binary_mask"] = c.image, c.mask, c.binary_mask
        g[p + "depth_map"], g[p + "fused_depth_map"] = c.depth_map, c.fused_depth_map
        if hasattr(c, "border_mask"):
            g[p + "border_mask"] = c.border_mask
        g[p + "near_far"] = np.array([c.near["bkg"], c.far["bkg"], c.near["human"], c.far["human"]])
        g[p + "frame"] = np.array([c.frame_id["frame_id"], c.frame_id["total_frames"]])
    return g


def record(ref, fn):
    """Runs fn() while recording the pixel lists that reach utils.ray_utils.shot_rays."""
    seen = []
    orig = ref.ray_utils.shot_rays

    def spy(cap, xys):
        seen.append(np.array(xys))
        return orig(cap, xys)
    ref.ray_utils.shot_rays = spy
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            out = fn()
    finally:
        ref.ray_utils.shot_rays = orig
    return out, seen


def batch_arrays(out, prefix):
    g = {}
    for k, v in out.items():
        g[prefix + k] = v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    return g


def main():
    import importlib.util
    ref = ref_import.load()

    def load_file(name):
        # the reference's datasets/ has no __init__.py and an installed `datasets` package shadows it: load by path
        spec = importlib.util.spec_from_file_location("ref_datasets_" + name, os.path.join(ref_import.REF_ROOT, "datasets", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    bg, hr = load_file("background_rays"), load_file("human_rays")
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        # ---------------- background batches (datasets/background_rays.py:41-139) ----------------
        caps = make_caps(ref, 3, seed=3)
        split = os.path.join(tmp, "train_split.txt")
        with open(split, "w") as f:
            f.write("\n".join(os.path.basename(c.image_path) for c in caps) + "\n")
        g.update(cap_arrays(caps, "bg_cap"))
        for tag, fused, nerft in (("bg", False, False), ("bgf", True, False), ("bgt", False, True)):
            opt = types.SimpleNamespace(rays_per_batch=500, use_fused_depth=fused, ablate_nerft=nerft)
            ds = bg.BackgroundRayDataset(opt, FakeScene(caps), "train", split)
            np.random.seed(7)
            random.seed(7)
            out, seen = record(ref, lambda: ds[0])
            bins = [len(s) for s in seen]
            assert sum(bins) == 500
            # which capture each recorded list belongs to: captures with a zero bin are skipped by the reference
            np.random.seed(7)
            draw = np.random.multinomial(500, np.ones(3) / 3.0)
            assert [int(b) for b in draw if b > 0] == bins
            g[tag + "_bins"] = draw
            g[tag + "_coords"] = np.concatenate(seen)
            g.update(batch_arrays(out, tag + "_out_"))

        # ---------------- human batches (datasets/human_rays.py:100-247) ----------------
        hcaps = make_caps(ref, 2, seed=5, human=True)
        body = synth_smpl.random_body(seed=1, center=(0.1, -0.05, -0.2))
        hsplit = os.path.join(tmp, "human_split.txt")
        with open(hsplit, "w") as f:
            f.write("\n".join(os.path.basename(c.image_path) for c in hcaps) + "\n")
        g.update(cap_arrays(hcaps, "hu_cap"))
        cache = {}
        for i, c in enumerate(hcaps):                       # data_io/cache_helper.py:16-36 with the reference's own function
            o, d = ref.ray_utils.shot_all_rays(c)
            near, far = ref.ray_utils.geometry_guided_near_far(torch.from_numpy(o).float(), torch.from_numpy(d).float(),
                                                               torch.from_numpy(body["verts"]).float(), body["geo_threshold"])
            nf = np.stack([near.numpy().reshape(c.shape[0], c.shape[1]), far.numpy().reshape(c.shape[0], c.shape[1])], -1)
            cache[os.path.basename(c.image_path)] = nf
            g[f"hu_cap{i}_cache"] = nf
        for tag, lpips, seed in (("hu", 0.0, 11), ("hup", 0.1, 12), ("hun", 0.1, 15)):
            opt = types.SimpleNamespace(rays_per_batch=1400, white_bkg=True, penalize_lpips=lpips, dilation=5,
                                        body_rays_ratio=0.6, border_rays_ratio=0.1, bkg_rays_ratio=0.3, geo_threshold=0.1,
                                        chunk=4096)
            ds = hr.HumanRayDataset(opt, FakeScene(hcaps), "train", hsplit, near_far_cache=cache)
            ds.cap_id = 1 if tag != "hu" else 0
            # the patch is sampled when random.random() < body_rays_ratio (:122): pick seeds that take each branch
            for s in range(seed, seed + 50):
                random.seed(s)
                np.random.seed(s)
                out, seen = record(ref, lambda: ds[0])
                want_patch = tag == "hup"
                if lpips == 0 or bool(int(out["patch_counter"])) == want_patch:
                    break
            assert lpips == 0 or bool(int(out["patch_counter"])) == want_patch
            g[tag + "_seg"] = np.array([len(x) for x in seen])
            g[tag + "_coords"] = np.concatenate(seen)
            g[tag + "_cap"] = np.array(ds.cap_id)
            g.update(batch_arrays(out, tag + "_out_"))
        g["hu_opt"] = np.array([1400, 5, 0.6, 0.1, 0.3])
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
