"""The C-ABI library loads without a GPU and exports every entry point include/hpvg.h declares; the ctypes binding
(hpvg/lib.py) covers exactly that set.  No compute call is made here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "hpvg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hpvg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from hpvg import lib
    names = _declared()
    assert len(names) >= 30
    handle = lib.load()
    for name in names:
        assert name in lib.PROTOTYPES, "include/hpvg.h declares %s but hpvg/lib.py does not bind it" % name
        getattr(handle, name)                      # raises AttributeError if the shared object does not export it
    for name in lib.PROTOTYPES:
        assert name in names, "hpvg/lib.py binds %s which include/hpvg.h does not declare" % name


def test_host_only_entry_points():
    from hpvg import lib
    handle = lib.load()
    assert handle.hpvg_version() >= 100
    assert lib.get_conv_backend() == lib.BACKEND_AUTO
    with pytest.raises(lib.HpvgError):
        lib.set_conv_backend(7)
    assert b"unknown backend" in handle.hpvg_last_error()
    lib.set_conv_backend(lib.BACKEND_AUTO)
    assert handle.hpvg_conv_wgrad_workspace(1, 64, 64, 4, 16, 16, 3, 1, 1, 1) > 0
    assert handle.hpvg_conv_wgrad_workspace(1, 8, 8, 4, 16, 16, 3, 1, 1, 1) == 0     # CUDA-core path needs none


def test_drop_in_modules_expose_the_reference_names():
    from modules import losses, networks_2d, networks_3d, utils
    for name in ("ConvBlock3D", "ConvBlock3DSN", "FeatureExtractor", "Encode3DVAE", "WDiscriminator3D", "GeneratorHPVAEGAN",
                 "GeneratorSG", "reparameterize", "weights_init", "get_activation"):
        assert hasattr(networks_3d, name), name
    for name in ("ConvBlock2D", "ConvBlock2DSN", "FeatureExtractor", "Encode2DVAE", "WDiscriminator2D", "GeneratorHPVAEGAN",
                 "reparameterize"):
        assert hasattr(networks_2d, name), name
    assert callable(losses.kl_criterion) and callable(utils.calc_gradient_penalty)
    assert utils.torch.__name__ == "torch"        # train_video.py:16 relies on `from modules.utils import *` exporting torch
    assert not hasattr(utils, "__all__")


def test_state_dict_keys_match_the_reference_contract(golden):
    """identical keys and shapes as the reference modules: the fixtures list the reference's own state"""
    from helpers import opt_from
    from modules import networks_3d
    fx = golden("hp3d_wide")
    g = networks_3d.GeneratorHPVAEGAN(opt_from(fx))
    for _ in range(fx["stages"]):
        g.init_next_stage()
    mine = [(k, tuple(v.shape)) for k, v in g.state_dict().items()]
    assert mine == [(k, tuple(s)) for k, s in fx["state"]]
    fd = golden("d3d_wide")
    d = networks_3d.WDiscriminator3D(opt_from(fd))
    assert [(k, tuple(v.shape)) for k, v in d.state_dict().items()] == [(k, tuple(s)) for k, s in fd["state"]]
