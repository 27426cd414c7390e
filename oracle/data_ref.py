"""ORACLE (test infrastructure only): numpy restatement of one sample of the reference's CPU data path, with the random
draws passed in, so the CUDA kernel `rcot_make_patches` can be compared bit for bit.

Follows /root/reference: util/image_utils.py:59-65 (crop_img), :133-163 (data_augmentation), :177-182
(random_augmentation draws mode 1..7), util/degradation_utils.py:21-27 (_add_gaussian_noise: float64 arithmetic,
clip, astype(uint8)), util/dataset_utils.py:215-278 (__getitem__: crop -> augmentation -> noise -> ToTensor).
Pinned against the reference's own functions in tests/test_data.py (CPU, where the reference is importable).
"""
import numpy as np


def crop_img(image, base=16):
    h, w = image.shape[0], image.shape[1]
    ch, cw = h % base, w % base
    return image[ch // 2:h - ch + ch // 2, cw // 2:w - cw + cw // 2, :]


def augment(image, mode):
    if mode == 0:
        return image
    if mode == 1:
        return np.flipud(image)
    out = np.rot90(image, k={2: 1, 3: 1, 4: 2, 5: 2, 6: 3, 7: 3}[mode])
    return np.flipud(out) if mode in (3, 5, 7) else out


def make_patch(clean_img, deg_img, y0, x0, P, mode, sigma, noise):
    """clean_img/deg_img: uint8 HWC (deg_img None for the denoise tasks); noise: (P, P, 3) standard normals.
    Returns (degraded, clean) float32 CHW in [0, 1] like torchvision's ToTensor."""
    c = crop_img(clean_img)[y0:y0 + P, x0:x0 + P]
    c = np.ascontiguousarray(augment(c, mode))
    if sigma > 0:
        d = np.clip(c + noise.astype(np.float64) * sigma, 0, 255).astype(np.uint8)
    else:
        d = np.ascontiguousarray(augment(crop_img(deg_img)[y0:y0 + P, x0:x0 + P], mode))
    to = lambda a: (a.transpose(2, 0, 1).astype(np.float32) / np.float32(255))
    return to(d), to(c)
