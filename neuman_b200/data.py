"""Ray batches for the background trainer produced on the device: host mirror of
`BackgroundRayDataset.__getitem__` (datasets/background_rays.py:41-139), SURVEY.md §8f-3.

The reference builds every batch on the CPU (np.argwhere over the full-resolution masks of every capture, fancy
indexing of the images, shot_rays in numpy) and ships it through a DataLoader; at a few milliseconds per training
step that producer would be the bottleneck.  Here the per-capture arrays (image, depth map, list of admissible
pixels) are uploaded once, and a batch is a handful of device gathers plus the ray-generation kernel.
Only the multinomial split of the batch over the captures stays on the host (np.random, as in the reference).
"""
import numpy as np
import torch

from . import ops


class BackgroundRayBatcher:
    """caps: capture objects with .image [H,W,3] uint8, .depth_map (and .fused_depth_map when
    opt.use_fused_depth), .mask / .binary_mask (/ .border_mask), .near['bkg'], .far['bkg'], .frame_id,
    .intrinsic_matrix, .cam_pose.camera_to_world, .shape -- what datasets/background_rays.py reads."""

    def __init__(self, opt, caps, device="cuda"):
        self.opt = opt
        self.batch_size = int(opt.rays_per_batch)
        self.device = torch.device(device)
        self.caps = list(caps)
        self.lut = torch.from_numpy((np.arange(256) / 255).astype(np.float32)).to(self.device)    # (img / 255).astype(float32), :78
        self.images, self.depths, self.valid = [], [], []
        for cap in self.caps:
            img = np.asarray(cap.image)
            self.images.append(torch.from_numpy(np.ascontiguousarray(img[..., :3])).to(self.device))
            dm = cap.fused_depth_map if getattr(opt, 'use_fused_depth', False) else cap.depth_map
            self.depths.append(torch.from_numpy(np.ascontiguousarray(dm, dtype=np.float32)).to(self.device))
            if getattr(opt, 'ablate_nerft', False):
                coords = None                                                   # whole image (:63-68)
            elif hasattr(cap, 'border_mask'):
                assert hasattr(cap, 'binary_mask')
                coords = np.argwhere((cap.border_mask | cap.mask) == 0)[:, ::-1]   # (x, y) (:69-73)
            elif hasattr(cap, 'binary_mask'):
                coords = np.argwhere(cap.mask == 0)[:, ::-1]                       # (:74-77)
            else:
                raise ValueError
            self.valid.append(None if coords is None else
                              torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int32)).to(self.device))

    def sample_coords(self, generator=None):
        """The random part of __getitem__: multinomial split over the captures (host RNG, :47), uniform pixels
        from each capture's admissible set (device RNG).  Returns a list of [num,2] int32 (x, y) tensors or None."""
        bins = np.random.multinomial(self.batch_size, np.ones(len(self.caps)) / float(len(self.caps)))
        out = []
        for cap, valid, num in zip(self.caps, self.valid, bins):
            if num == 0:
                out.append(None)
                continue
            if valid is None:
                h, w = cap.image.shape[:2]
                y = torch.randint(0, h, (int(num),), device=self.device, generator=generator)
                x = torch.randint(0, w, (int(num),), device=self.device, generator=generator)
                out.append(torch.stack([x, y], 1).int())
            else:
                idx = torch.randint(0, valid.shape[0], (int(num),), device=self.device, generator=generator)
                out.append(valid[idx])
        return out

    def batch_from_coords(self, coords_list):
        """The deterministic part (:78-139) for given pixels: same keys, shapes and dtypes as the reference batch
        (without the DataLoader's leading axis), CUDA tensors."""
        cols, deps, origs, dirs, nears, fars, bkg, viewf = [], [], [], [], [], [], [], []
        for cap, img, dm, xy in zip(self.caps, self.images, self.depths, coords_list):
            if xy is None or xy.shape[0] == 0:
                continue
            x, y = xy[:, 0].long(), xy[:, 1].long()
            cols.append(self.lut[img[y, x].long()])
            deps.append(dm[y, x])
            o, d = ops.shot_rays(cap, xy)
            origs.append(o)
            dirs.append(d)
            num = xy.shape[0]
            nears.append(torch.full((num, 1), float(cap.near['bkg']), device=self.device))
            fars.append(torch.full((num, 1), float(cap.far['bkg']), device=self.device))
            bkg.append(torch.ones(num, 1, dtype=torch.long, device=self.device))
            fid = getattr(cap, 'frame_id', {'frame_id': 0, 'total_frames': 1})
            viewf.append(torch.full((num, 1), float(np.float32(fid['frame_id'] / fid['total_frames'])), device=self.device))
        return {'color': torch.cat(cols), 'depth': torch.cat(deps), 'origin': torch.cat(origs), 'direction': torch.cat(dirs),
                'near': torch.cat(nears), 'far': torch.cat(fars), 'is_bkg': torch.cat(bkg), 'viewf_list': torch.cat(viewf)}

    def __call__(self, generator=None):
        return self.batch_from_coords(self.sample_coords(generator))


def near_far_cache_device(cap, verts, geo_threshold=ops.DEFAULT_GEO_THRESH, device="cuda"):
    """export_near_far_cache (data_io/cache_helper.py:16-36) kept on the device: [H,W,2] float32 (near, far) of
    geometry_guided_near_far for every pixel (near = inf, far = -inf where the ray misses)."""
    o, d = ops.shot_all_rays(cap, device=device, mode=0)
    near, far = ops.geometry_guided_near_far(o, d, ops._f32(verts, o.device), geo_threshold)
    H, W = int(cap.shape[0]), int(cap.shape[1])
    return torch.stack([near.reshape(H, W), far.reshape(H, W)], -1)


PATCH_SIZE = 32                      # utils/constant.py:8
PATCH_SIZE_SQUARED = PATCH_SIZE ** 2


def get_left_upper_corner(h, w, pos, size=PATCH_SIZE):
    """datasets/human_rays.py:18-34: left upper corner (x, y) of a size x size patch centred (as centred as the image
    allows) at pos = (x, y)."""
    lu_x = min(max(int(pos[0]) - size // 2, 0), w - size)
    lu_y = min(max(int(pos[1]) - size // 2, 0), h - size)
    return lu_x, lu_y


class HumanRayBatcher:
    """`HumanRayDataset.__getitem__` (datasets/human_rays.py:100-247) on the device, including the 32x32 patch branch
    that the LPIPS term of the human trainer needs (opt.penalize_lpips > 0, :114-126, :163-183; the LPIPS network itself is
    the caller's).  caps: captures with .image, .mask, .binary_mask, .border_mask (when opt.dilation > 0), .near/.far
    {'bkg','human'}, .frame_id and the camera; near_far: one [H,W,>=2] array/tensor per capture (near_far_cache_device, or
    the reference's .npy cache).

    A batch is a list of segments (ray_key, pixels) in the reference's order: with a patch, the first PATCH_SIZE_SQUARED
    rays are the patch in row-major order (or, when the coin of :122 falls the other way, a body/border/background split of
    the same size), followed by the split of the remaining rays."""

    KEYS = ('num_body_rays', 'num_border_rays', 'num_bkg_rays')

    def __init__(self, opt, caps, near_far, device="cuda"):
        self.opt, self.batch_size, self.device = opt, int(opt.rays_per_batch), torch.device(device)
        self.num_patch = 1 if getattr(opt, 'penalize_lpips', 0) > 0 else 0          # (:72)
        if self.num_patch:
            assert self.batch_size > PATCH_SIZE_SQUARED                               # (:117)
        self.caps = list(caps)
        self.lut = torch.from_numpy((np.arange(256) / 255).astype(np.float32)).to(self.device)
        self.images, self.binary, self.cache, self.sets = [], [], [], []

        def up(a, dtype):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self.device)
        for cap, nf in zip(self.caps, near_far):
            self.images.append(up(np.asarray(cap.image)[..., :3], np.uint8))
            self.binary.append(up(cap.binary_mask, np.int64))
            nf = nf if isinstance(nf, torch.Tensor) else torch.from_numpy(np.asarray(nf))
            self.cache.append(nf[..., :2].to(self.device).float().contiguous())
            sets = {'num_body_rays': up(np.argwhere(cap.mask != 0)[:, ::-1], np.int32),          # (:156)
                    'num_bkg_rays': up(np.argwhere(cap.mask == 0)[:, ::-1], np.int32)}            # (:160)
            if getattr(opt, 'dilation', 0) > 0:
                sets['num_border_rays'] = up(np.argwhere(cap.border_mask == 1)[:, ::-1], np.int32)   # (:158)
            self.sets.append(sets)

    def get_num_rays_dict(self, num):
        """datasets/human_rays.py:81-97."""
        o = self.opt
        arr = np.array([int(round(num * o.body_rays_ratio)),
                        int(round(num * o.border_rays_ratio)) if o.dilation > 0 else 0,
                        int(round(num * o.bkg_rays_ratio))])
        arr[arr.argmax()] += num - arr.sum()
        assert arr.min() >= 0 and arr.sum() == num
        return dict(zip(self.KEYS, (int(a) for a in arr)))

    def plan(self, need_patch):
        """The (ray_key, count) segments of one batch (:112-126, :145-152)."""
        bins = [self.batch_size] if self.num_patch == 0 else [PATCH_SIZE_SQUARED, self.batch_size - PATCH_SIZE_SQUARED]
        segs = []
        for i, num in enumerate(bins):
            if num == 0:
                continue
            if self.num_patch == 1 and need_patch and i == 0:
                segs.append(('num_patch_rays', num))
            else:
                segs += [(k, n) for k, n in self.get_num_rays_dict(num).items() if n > 0]
        return segs

    def patch_coords(self, cap_index, seed_xy):
        """The PATCH_SIZE x PATCH_SIZE pixels around seed_xy = (x, y) in row-major order as [n,2] (x, y) (:163-183)."""
        h, w = self.images[cap_index].shape[:2]
        lu_x, lu_y = get_left_upper_corner(h, w, seed_xy)
        ys = torch.arange(lu_y, lu_y + PATCH_SIZE, device=self.device, dtype=torch.int32)
        xs = torch.arange(lu_x, lu_x + PATCH_SIZE, device=self.device, dtype=torch.int32)
        return torch.stack([xs[None, :].expand(PATCH_SIZE, -1), ys[:, None].expand(-1, PATCH_SIZE)], -1).reshape(-1, 2).contiguous()

    def sample_coords(self, cap_index, generator=None, need_patch=None):
        """The random part: returns [(ray_key, [num,2] (x, y) int32), ...]."""
        if need_patch is None:
            need_patch = bool(np.random.random() < self.opt.body_rays_ratio)        # random.random() < body_rays_ratio (:122)
        out = []
        for key, num in self.plan(need_patch):
            if key == 'num_patch_rays':
                pool = self.sets[cap_index]['num_body_rays']                        # random.choice(argwhere(mask != 0)) (:163)
                seed = pool[torch.randint(0, pool.shape[0], (1,), device=self.device, generator=generator)][0].tolist()
                out.append((key, self.patch_coords(cap_index, seed)))
            else:
                pool = self.sets[cap_index][key]
                out.append((key, pool[torch.randint(0, pool.shape[0], (num,), device=self.device, generator=generator)]))
        return out

    def batch_from_coords(self, cap_index, coords):
        """The deterministic part (:176-247) for given pixels: [(ray_key, [num,2] (x, y)), ...] in batch order (or a
        {ray_key: pixels} dict, taken in KEYS order)."""
        if isinstance(coords, dict):
            coords = [(k, coords[k]) for k in self.KEYS if coords.get(k) is not None]
        cap, img, cache = self.caps[cap_index], self.images[cap_index], self.cache[cap_index]
        cols, origs, dirs, hn, hf, bn, bf, isb, hit = [], [], [], [], [], [], [], [], []
        patch_counter = 0
        for key, xy in coords:
            if xy is None or xy.shape[0] == 0:
                continue
            patch_counter += int(key == 'num_patch_rays')
            x, y = xy[:, 0].long(), xy[:, 1].long()
            num = xy.shape[0]
            cols.append(self.lut[img[y, x].long()])
            isb.append(1 - self.binary[cap_index][y, x])
            o, d = ops.shot_rays(cap, xy)
            origs.append(o)
            dirs.append(d)
            c = cache[y, x]
            valid = c[:, 0] < c[:, 1]
            hn.append(torch.where(valid, c[:, 0], torch.full_like(c[:, 0], float(cap.near['human'])))[:, None])
            hf.append(torch.where(valid, c[:, 1], torch.full_like(c[:, 1], float(cap.far['human'])))[:, None])
            bn.append(torch.full((num, 1), float(cap.near['bkg']), device=self.device))
            bf.append(torch.full((num, 1), float(cap.far['bkg']), device=self.device))
            hit.append(valid.long())
        fid = cap.frame_id
        return {'color': torch.cat(cols), 'origin': torch.cat(origs), 'direction': torch.cat(dirs),
                'human_near': torch.cat(hn), 'human_far': torch.cat(hf), 'bkg_near': torch.cat(bn), 'bkg_far': torch.cat(bf),
                'is_bkg': torch.cat(isb), 'is_hit': torch.cat(hit),
                'cur_view_f': fid['frame_id'] / fid['total_frames'], 'cur_view': fid['frame_id'], 'cap_id': cap_index,
                'patch_counter': torch.tensor(patch_counter)}

    def __call__(self, cap_index=None, generator=None, need_patch=None):
        if cap_index is None:
            cap_index = int(np.random.randint(len(self.caps)))          # random.choice(self.inclusions) (:105)
        return self.batch_from_coords(cap_index, self.sample_coords(cap_index, generator, need_patch))
