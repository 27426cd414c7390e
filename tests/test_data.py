"""Data path.  CPU: the numpy oracle (oracle/data_ref.py) against the reference's OWN functions
(crop_img, random_augmentation, Degradation._add_gaussian_noise) under the same seeds.  GPU: `rcot_make_patches`
against the oracle, bit for bit, for every augmentation mode, ragged image sizes and both sample kinds."""
import random

import numpy as np
import pytest
import torch


def _ref_utils():
    import importlib
    import sys
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference files not present")
    ref = ref_shim.ref_dir()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "util" or k.startswith("util.")}
    sys.path.insert(0, ref)
    try:
        iu = importlib.import_module("util.image_utils")
        du = importlib.import_module("util.degradation_utils")
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "util" or k.startswith("util.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    return iu, du


def test_oracle_matches_reference_functions():
    from oracle import data_ref
    iu, du = _ref_utils()
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (75, 101, 3)).astype(np.uint8)
    deg = rs.randint(0, 256, (75, 101, 3)).astype(np.uint8)
    assert np.array_equal(iu.crop_img(img, base=16), data_ref.crop_img(img))
    P = 32
    for seed in range(16):
        random.seed(seed)
        c0 = data_ref.crop_img(img)[3:3 + P, 7:7 + P]
        d0 = data_ref.crop_img(deg)[3:3 + P, 7:7 + P]
        d_ref, c_ref = iu.random_augmentation(d0, c0)              # draws mode = random.randint(1, 7)
        random.seed(seed)
        mode = random.randint(1, 7)
        d_o, c_o = data_ref.make_patch(img, deg, 3, 7, P, mode, 0.0, None)
        assert np.array_equal(c_o, c_ref.transpose(2, 0, 1).astype(np.float32) / np.float32(255))
        assert np.array_equal(d_o, d_ref.transpose(2, 0, 1).astype(np.float32) / np.float32(255))
        # denoise branch: noise = np.random.randn(*patch.shape) inside the reference's Degradation
        class A:
            patch_size = P
        D = du.Degradation(A())
        np.random.seed(seed)
        noisy_ref, _ = D._add_gaussian_noise(c_ref, sigma=25)
        np.random.seed(seed)
        noise = np.random.randn(P, P, 3)
        d_o, _ = data_ref.make_patch(img, None, 3, 7, P, mode, 25.0, noise)
        assert np.array_equal(d_o, noisy_ref.transpose(2, 0, 1).astype(np.float32) / np.float32(255))


@pytest.mark.gpu
def test_make_patches_bit_exact(cuda_lib):
    from oracle import data_ref
    from rcot_b200.data import DevicePool, DeviceTrainData
    rs = np.random.RandomState(1)
    pool = DevicePool("cuda")
    imgs = []
    for (h, w) in [(75, 101), (64, 64), (130, 97), (48, 200)]:
        c = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        d = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        imgs.append((c, d, pool.add(c), pool.add(d)))
    pool.finalize()
    P = 32
    data = DeviceTrainData(pool, [], P)
    draws, want = [], []
    noise = torch.randn(32, P, P, 3, generator=torch.Generator().manual_seed(0))
    k = 0
    for mode in range(8):
        for (c, d, ci, di) in imgs:
            Hc, Wc = c.shape[0] - c.shape[0] % 16, c.shape[1] - c.shape[1] % 16
            y0, x0 = rs.randint(0, Hc - P + 1), rs.randint(0, Wc - P + 1)
            de_id = (0, 1, 2, 3, 4, 7)[k % 6]
            draws.append((de_id, ci, di if de_id >= 3 else None, y0, x0, mode))
            sig = {0: 15.0, 1: 25.0, 2: 50.0}.get(de_id, 0.0)
            want.append(data_ref.make_patch(c, d, y0, x0, P, mode, sig, noise[k].numpy()))
            k += 1
    ids, deg, cln = data.assemble(draws, noise=noise.cuda())
    assert ids.tolist() == [d[0] for d in draws]
    for i, (d_o, c_o) in enumerate(want):
        assert np.array_equal(cln[i].cpu().numpy(), c_o), ("clean", i, draws[i])
        assert np.array_equal(deg[i].cpu().numpy(), d_o), ("degraded", i, draws[i])
    # the draw() path: in-range crops, modes 1..7, device-side noise
    from rcot_b200.data import synthetic_pool
    pool2, samples = synthetic_pool(6, 80, 112, ["denoise_25", "derain", "dehaze"], device="cuda")
    dd = DeviceTrainData(pool2, samples, 64, seed=3)
    ids, deg, cln = dd.batch(16)
    assert deg.shape == (16, 3, 64, 64) and 0 <= deg.min() and deg.max() <= 1 and set(ids.tolist()) <= {1, 3, 4}
    q = cln * 255
    assert torch.equal(q.round(), torch.floor(q + 0.5)) and (q - q.round()).abs().max() < 1e-4   # uint8 grid
