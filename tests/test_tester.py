"""tester.py (reference tester.py:54-112): a pickled checkpoint -> folder inference -> PNGs + PSNR, whole-image and
tiled, on images whose sides are not multiples of 8."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tester_runs_on_checkpoint(cuda_lib, tmp_path):
    import numpy as np
    from PIL import Image

    import Net_Restormer as N
    import tester
    torch.manual_seed(0)
    T = N.T_net(decoder=True).cuda()
    ck = tmp_path / "m.pth"
    torch.save({"epoch": 1, "Tnet": T, "Fnet": None}, ck)
    deg_d, tar_d = tmp_path / "deg", tmp_path / "tar"
    deg_d.mkdir(); tar_d.mkdir()
    rng = np.random.default_rng(0)
    for i, (h, w) in enumerate([(70, 52), (64, 96)]):
        tar = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        deg = np.clip(tar.astype(np.int32) + rng.integers(-20, 20, (h, w, 3)), 0, 255).astype(np.uint8)
        Image.fromarray(tar).save(tar_d / f"{i}.png")
        Image.fromarray(deg).save(deg_d / f"{i}.png")
    args = ["--model", str(ck), "--degset", str(deg_d) + "/", "--tarset", str(tar_d) + "/", "--save", str(tmp_path / "o") + "/",
            "--savetar", str(tmp_path / "t") + "/", "--saveres", str(tmp_path / "r") + "/"]
    ps = tester.main(args)
    assert len(ps) == 2 and all(p == p for p in ps)
    out0 = np.array(Image.open(tmp_path / "o" / "0.png"))
    assert out0.shape == (64, 48, 3)                       # cropped to multiples of 8
    # the PNG equals a direct forward of the same (cropped) image
    deg0 = np.array(Image.open(deg_d / "0.png").convert("RGB"))[:64, :48]
    x = torch.from_numpy(deg0.transpose(2, 0, 1).copy()).float().div(255).unsqueeze(0).cuda()
    with torch.no_grad():
        y = T(x).clamp(0, 1).mul(255).add(0.5).clamp(0, 255).byte().squeeze(0).permute(1, 2, 0).cpu().numpy()
    assert np.abs(y.astype(int) - out0.astype(int)).max() <= 1
    ps_t = tester.main(args + ["--tile", "48", "--tile_overlap", "16"])
    assert len(ps_t) == 2 and all(p == p for p in ps_t)
