"""Device-side training data: the reference's TrainDataset semantics -- centre crop to multiples of 16,
random P x P crop, one of 7 flip/rot90 augmentations, uint8-grid Gaussian noise for the denoise tasks
(util/dataset_utils.py:215-278, util/image_utils.py:59-65,133-182, util/degradation_utils.py:21-27) -- for a whole batch
in ONE kernel launch (`rcot_make_patches`) from uint8 images resident in HBM, instead of a per-item PIL/numpy pipeline
behind a DataLoader (which cannot feed > 1 k images/s with the reference's default of 0 workers).

The host only draws the per-sample integers (image id, crop origin, mode) from Python's `random` like the reference
(`random.randint(1, 7)`, util/image_utils.py:179); the Gaussian noise is drawn on the device (torch CUDA generator).
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from . import _lib

SIGMAS = {0: 15.0, 1: 25.0, 2: 50.0}          # de_id -> sigma (util/degradation_utils.py:30-38)


class PatchDesc(C.Structure):
    _fields_ = [("clean_off", C.c_int64), ("deg_off", C.c_int64), ("H", C.c_int32), ("W", C.c_int32),
                ("y0", C.c_int32), ("x0", C.c_int32), ("mode", C.c_int32), ("sigma", C.c_float)]


class DevicePool:
    """uint8 HWC images packed into one device buffer.  add() returns the image's index."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.host, self.meta, self.size = [], [], 0
        self.buf = None

    def add(self, img_u8: np.ndarray) -> int:
        assert img_u8.dtype == np.uint8 and img_u8.ndim == 3 and img_u8.shape[2] == 3
        a = np.ascontiguousarray(img_u8)
        self.host.append(a)
        self.meta.append((self.size, a.shape[0], a.shape[1]))
        self.size += (a.size + 15) // 16 * 16
        self.buf = None
        return len(self.meta) - 1

    def finalize(self):
        flat = np.zeros(self.size, dtype=np.uint8)
        for a, (off, _, _) in zip(self.host, self.meta):
            flat[off:off + a.size] = a.reshape(-1)
        self.buf = torch.from_numpy(flat).to(self.device)
        self.host = []
        return self


class DeviceTrainData:
    """samples: list of (de_id, clean_index, degraded_index or None) over a finalized DevicePool."""

    def __init__(self, pool: DevicePool, samples, patch, seed=0):
        self.pool, self.samples, self.P = pool, samples, patch
        self.rng = random.Random(seed)
        self.gen = torch.Generator(device=pool.device).manual_seed(seed)

    def draw(self, B):
        """The per-sample integers the reference draws on the CPU (sample order: uniform with replacement)."""
        out = []
        for _ in range(B):
            de_id, ci, di = self.samples[self.rng.randrange(len(self.samples))]
            _, H, W = self.pool.meta[ci]
            Hc, Wc = H - H % 16, W - W % 16
            y0, x0 = self.rng.randint(0, Hc - self.P), self.rng.randint(0, Wc - self.P)
            out.append((de_id, ci, di, y0, x0, self.rng.randint(1, 7)))
        return out

    def assemble(self, draws, noise=None):
        """draws: [(de_id, clean_idx, deg_idx, y0, x0, mode)]; noise: optional [B,P,P,3] float32 device tensor.
        Returns (de_id int64 [B] (CPU), degraded, clean) -- the batch layout of the reference's loader."""
        if self.pool.buf is None:
            self.pool.finalize()
        B, P = len(draws), self.P
        arr = (PatchDesc * B)()
        need_noise = False
        for i, (de_id, ci, di, y0, x0, mode) in enumerate(draws):
            off, H, W = self.pool.meta[ci]
            sig = SIGMAS.get(de_id, 0.0)
            need_noise |= sig > 0
            arr[i] = PatchDesc(off, self.pool.meta[di][0] if di is not None else 0, H, W, y0, x0, mode, sig)
        dev = self.pool.device
        desc = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        if need_noise and noise is None:
            noise = torch.randn(B, P, P, 3, device=dev, generator=self.gen)
        deg = torch.empty(B, 3, P, P, device=dev)
        cln = torch.empty(B, 3, P, P, device=dev)
        lib = _lib.lib()
        _lib.check(lib.rcot_make_patches(C.c_void_p(self.pool.buf.data_ptr()), C.c_void_p(desc.data_ptr()),
                                         C.c_void_p(noise.data_ptr() if noise is not None else 0),
                                         C.c_void_p(deg.data_ptr()), C.c_void_p(cln.data_ptr()), B, P,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "make_patches")
        self._keep = (desc, noise)      # alive until the stream has consumed them
        return torch.tensor([d[0] for d in draws]), deg, cln

    def batch(self, B):
        return self.assemble(self.draw(B))


def synthetic_pool(n_images, H, W, de_types, seed=0, device="cuda"):
    """A pool of seeded synthetic uint8 images + the sample list for `de_types` (ids as util/dataset_utils.py:40):
    denoise_*: clean only; derain: sparse bright streaks; dehaze: t*clean + A*(1-t); others: noisy copy."""
    ids = {'denoise_15': 0, 'denoise_25': 1, 'denoise_50': 2, 'derain': 3, 'dehaze': 4, 'deblur': 5, 'lowlight': 6,
           'single': 7}
    rs = np.random.RandomState(seed)
    pool, samples = DevicePool(device), []
    for i in range(n_images):
        clean = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        ci = pool.add(clean)
        de_id = ids[de_types[i % len(de_types)]]
        if de_id < 3:
            samples.append((de_id, ci, None))
            continue
        c = clean.astype(np.float64)
        if de_id == 3:
            d = c + (rs.rand(H, W, 1) > 0.97) * 255 * (0.4 + 0.4 * rs.rand())
        elif de_id == 4:
            t, A = 0.3 + 0.6 * rs.rand(), 0.7 + 0.3 * rs.rand()
            d = c * t + 255 * A * (1 - t)
        else:
            d = c + 25 * rs.randn(H, W, 3)
        samples.append((de_id, ci, pool.add(np.clip(d, 0, 255).astype(np.uint8))))
    return pool.finalize(), samples
