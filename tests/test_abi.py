"""CPU checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol that
include/rcot_b200.h declares (no compute call is made: there is no GPU here), and the Python host
surface mirrors the reference's (CLI flags, class names)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from rcot_b200 import _lib
    path = _lib.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "rcot_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(rcot_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/rcot_b200.h but not exported: {missing}"
    assert lib.rcot_version() >= 100
    lib.rcot_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.rcot_last_error(), bytes)


def test_argument_errors_are_reported_not_raised_across_the_abi():
    """NULL / bad sizes return a negative code and set rcot_last_error (no GPU work is launched)."""
    from rcot_b200 import _lib
    lib = ctypes.CDLL(_lib.build())
    lib.rcot_last_error.restype = ctypes.c_char_p
    assert lib.rcot_pm_gemm(None, None) < 0
    assert b"null" in lib.rcot_last_error()
    assert lib.rcot_ln_stats(None, ctypes.c_int64(0), 1, 1, 1, None, None) < 0
    assert lib.rcot_rmsprop(None, None, None, ctypes.c_int64(0), ctypes.c_float(0), ctypes.c_float(0),
                            ctypes.c_float(0), ctypes.c_float(1), None) < 0


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the parameter structs have the sizes the C compiler gives them."""
    import subprocess
    import tempfile
    from rcot_b200 import ops
    src = '#include "rcot_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(rcot_pack_desc),' \
          ' sizeof(rcot_pm_params), sizeof(rcot_pk_params), sizeof(rcot_dw_params), sizeof(rcot_attn_params)); return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    got = [ctypes.sizeof(t) for t in (ops.PackDesc, ops.PMParams, ops.PKParams, ops.DWParams, ops.AttnParams)]
    assert got == sizes, (got, sizes)


def test_cli_surface_matches_reference():
    import trainer
    flags = {a.option_strings[0] for a in trainer.parser._actions if a.option_strings}
    ref_flags = {"--batchSize", "--nEpochs", "--lr", "--step", "--cuda", "--resume", "--start-epoch", "--threads",
                 "--pretrained", "--gpus", "--pairnum", "--de_type", "--denoise_dir", "--derain_dir", "--dehaze_dir",
                 "--degset", "--tarset", "--Sigma", "--sigma", "--optimizer", "--type", "--patch_size", "--num_workers",
                 "--data_file_dir"}
    assert ref_flags <= flags
    d = trainer.parser.parse_args([])
    assert (d.batchSize, d.nEpochs, d.lr, d.step, d.pairnum, d.Sigma, d.sigma, d.optimizer, d.patch_size) == \
        (4, 200, 1e-4, 20, 0, 10000, 1, "RMSprop", 64)
    trainer.opt = d
    assert trainer.adjust_learning_rate(None, 0) == 1e-4 and abs(trainer.adjust_learning_rate(None, 20) - 1e-5) < 1e-12


def test_synthetic_dataset_item_layout():
    import torch
    import trainer
    ds = trainer.SyntheticPairs(6, 32, ["denoise_25", "derain", "dehaze"], seed=3)
    (name, de_id), deg, clean = ds[1]
    assert de_id == 3 and deg.shape == clean.shape == (3, 32, 32) and deg.dtype == torch.float32
    assert 0 <= deg.min() and deg.max() <= 1
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=3)))
    assert batch[0][1].tolist() == [1, 3, 4] and batch[1].shape == (3, 3, 32, 32)


def test_checkpoint_pickle_layout(tmp_path):
    """{"epoch","Tnet","Fnet"} whole-module pickles resolve to Net_Restormer.* classes and carry no kernel state."""
    import torch
    import Net_Restormer as N
    F = N.F_net(patch_size=32)
    blk = N.TransformerBlock(48, 1, 2.66, False, 'WithBias')
    path = tmp_path / "ck.pth"
    torch.save({"epoch": 3, "Tnet": blk, "Fnet": F}, path)
    ck = torch.load(path, weights_only=False)
    assert ck["epoch"] == 3 and type(ck["Fnet"]).__module__ == "Net_Restormer"
    assert "_program" not in ck["Fnet"].__dict__
    assert list(ck["Fnet"].state_dict()) == list(F.state_dict())
