"""Drop-in for the reference's ``tester.py`` (inference over a folder pair, /root/reference/tester.py:13-124): same flags
and defaults, same checkpoint contract (``torch.load(opt.model)["Tnet"]`` is a callable pickled module, :54 -- files
written by either code base work), same outputs (restored / target / 2*residual PNGs, :110-112) -- with T_net running on
the sm_100a kernels.

Differences, all stated:  * images are cropped to multiples of 8 (the net has three 2x down-samplings; the reference
crops to multiples of 4, :77-84, and then fails inside the net for h % 8 == 4);  * PSNR is computed here (the
reference's evaluate.py needs skimage, and FID needs downloaded Inception weights: both are outside the hot path and
only run when those packages / weights are importable);  * additive ``--tile N`` runs overlapping N x N tiles with
linear blending for images that do not fit whole (an approximation: MDTA's channel attention is global over pixels).
"""
from __future__ import annotations

import argparse
import glob
import math
import os

import numpy as np
import torch

parser = argparse.ArgumentParser(description="RCOT tester (B200-native T_net)")
parser.add_argument("--cuda", action="store_true", help="use cuda? (always on, as in the reference)")
parser.add_argument("--model", default="./checkpoint/model_Dehazing__99_10.0.pth", type=str, help="model path")
parser.add_argument("--degset", default="./datasets/Dehazing/outdoor/hazy/", type=str, help="degraded data")
parser.add_argument("--tarset", default="./datasets/Dehazing/outdoor/gt/", type=str, help="target data")
parser.add_argument("--saveres", default="./results/Dehazing/RES/", type=str, help="savepath, Default: residual")
parser.add_argument("--save", default="./results/Dehazing/OUT/", type=str, help="savepath, Default: results")
parser.add_argument("--savetar", default="./results/Dehazing/TAR/", type=str, help="savepath, Default: targets")
parser.add_argument("--gpus", default="0", type=str, help="gpu ids")
# ---- additive
parser.add_argument("--tile", default=0, type=int, help="tile size (multiple of 8) for tiled inference; 0 = whole image")
parser.add_argument("--tile_overlap", default=32, type=int)


def PSNR(pred, gt, shave_border=0):
    height, width = pred.shape[:2]
    pred = pred[shave_border:height - shave_border, shave_border:width - shave_border]
    gt = gt[shave_border:height - shave_border, shave_border:width - shave_border]
    rmse = math.sqrt(((pred - gt) ** 2).mean())
    return 100 if rmse == 0 else 20 * math.log10(1.0 / rmse)


def tiled_forward(Tnet, x, tile, overlap):
    """Overlapping tiles blended with a linear ramp (tile, overlap multiples of 8)."""
    _, _, H, W = x.shape
    if tile <= 0 or (H <= tile and W <= tile):
        return Tnet(x)
    step = tile - overlap
    out = torch.zeros_like(x)
    wsum = torch.zeros(1, 1, H, W, device=x.device)
    ramp = torch.minimum(torch.arange(1, tile + 1), torch.arange(tile, 0, -1)).float().clamp(max=max(overlap, 1))
    ys = sorted({min(y, max(H - tile, 0)) for y in range(0, max(H - overlap, 1), step)})
    xs = sorted({min(x0, max(W - tile, 0)) for x0 in range(0, max(W - overlap, 1), step)})
    for y in ys:
        for x0 in xs:
            th, tw = min(tile, H - y), min(tile, W - x0)
            w2 = (ramp[:th].view(-1, 1) * ramp[:tw].view(1, -1)).to(x.device).view(1, 1, th, tw)
            o = Tnet(x[:, :, y:y + th, x0:x0 + tw].contiguous())
            out[:, :, y:y + th, x0:x0 + tw] += o * w2
            wsum[:, :, y:y + th, x0:x0 + tw] += w2
    return out / wsum


def main(argv=None):
    from PIL import Image
    from torchvision.utils import save_image
    opt = parser.parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", str(opt.gpus))
    if not torch.cuda.is_available():
        raise Exception("No GPU found (rcot_b200 has no CPU path)")
    for d in (opt.save, opt.savetar, opt.saveres):
        os.makedirs(d, exist_ok=True)
    Tnet = torch.load(opt.model, weights_only=False)["Tnet"].cuda()
    deg_list, tar_list = sorted(glob.glob(opt.degset + "*")), sorted(glob.glob(opt.tarset + "*"))
    psnrs = []
    with torch.no_grad():
        for deg_name, tar_name in zip(deg_list, tar_list):
            name = os.path.basename(tar_name)
            print("Processing ", deg_name)
            deg_img = np.array(Image.open(deg_name).convert('RGB'))
            tar_img = np.array(Image.open(tar_name).convert('RGB'))
            if deg_img.shape != tar_img.shape:
                continue
            h, w = deg_img.shape[0] // 8 * 8, deg_img.shape[1] // 8 * 8
            if h == 0 or w == 0:
                continue
            deg = torch.from_numpy(deg_img[:h, :w].transpose(2, 0, 1).copy()).float().div(255).unsqueeze(0).cuda()
            tar = torch.from_numpy(tar_img[:h, :w].transpose(2, 0, 1).copy()).float().div(255).unsqueeze(0)
            out = tiled_forward(Tnet, deg, opt.tile, opt.tile_overlap)
            res = (deg - out).cpu()
            out = out.cpu()
            save_image(res * 2, os.path.join(opt.saveres, name))
            save_image(out, os.path.join(opt.save, name))
            save_image(tar, os.path.join(opt.savetar, name))
            psnrs.append(PSNR(out.clamp(0, 1).squeeze(0).numpy(), tar.squeeze(0).numpy()))
    if psnrs:
        print("PSNR: Averyge {:.5f},   best {:.5f},   worst {:.5f}".format(sum(psnrs) / len(psnrs), max(psnrs), min(psnrs)))
    try:                                    # out of the hot path: only when the reference's metric stack is importable
        import fid_score                    # noqa: F401  (needs pytorch_fid + downloaded Inception weights)
        print('FID value:', fid_score.calculate_fid_given_paths([opt.savetar, opt.save], batch_size=50, device='cuda',
                                                                 dims=2048, num_workers=0))
    except Exception as e:
        print(f"FID / SSIM skipped ({type(e).__name__}): metric stack not available offline")
    return psnrs


if __name__ == "__main__":
    main()
