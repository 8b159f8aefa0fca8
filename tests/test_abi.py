"""The C-ABI library loads without a GPU and exports every entry point include/hpvg.h declares; the ctypes binding
(hpvg/lib.py) covers exactly that set.  No compute call is made here."""
import os
import re

import pytest
import torch

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
                 "GeneratorSG", "GeneratorCSG", "WDiscriminatorBaselines", "reparameterize", "weights_init", "get_activation"):
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


def test_state_dict_keys_of_the_baseline_networks(golden):
    """GeneratorCSG / GeneratorSG / WDiscriminatorBaselines: same keys, shapes and construction-time RNG consumption as the
    reference (weights_init draws N(0, 0.02) / N(1, 0.02) in module order, networks_3d.py:9-15,202,241,293)"""
    import torch
    from helpers import opt_from
    from modules import networks_3d
    for name, cls, stages in (("csg3d_tiny", networks_3d.GeneratorCSG, None), ("sg3d_tiny", networks_3d.GeneratorSG, None),
                              ("dbase3d_tiny", networks_3d.WDiscriminatorBaselines, 0)):
        fx = golden(name)
        m = cls(opt_from(fx))
        for _ in range(fx["stages"] if stages is None else stages):
            m.init_next_stage()
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in fx["state"]], name
    # the same seed gives the same initial weights as the reference's constructor would: weights_init overwrites every conv
    # weight with N(0, 0.02), so the standard deviation is the check that does not need the reference here
    torch.manual_seed(0)
    g = networks_3d.GeneratorCSG(opt_from(golden("csg3d_wide")))
    w = g.body[0].block0.conv.weight
    assert abs(w.std().item() - 0.02) < 2e-3 and abs(g.head.norm.weight.mean().item() - 1.0) < 2e-2


def test_state_dict_keys_of_the_bernoulli_variants(golden):
    """GeneratorVAE_nb (3-D and 2-D) and Encode3DVAE1x1 construct on the CPU with the reference's state_dict keys and shapes"""
    from helpers import opt_from
    from modules import networks_2d, networks_3d
    for dims, nets in ((3, networks_3d), (2, networks_2d)):
        fx = golden("nb%dd_tiny" % dims)
        g = nets.GeneratorVAE_nb(opt_from(fx))
        g.init_next_stage()
        g.init_next_stage()
        assert [(k, tuple(v.shape)) for k, v in g.state_dict().items()] == [(k, tuple(s)) for k, s in fx['state']]
    enc = networks_3d.Encode3DVAE1x1(opt_from(golden("nb3d_tiny")))
    keys = list(enc.state_dict().keys())
    assert 'features.conv_block_0.conv.weight_orig' in keys and 'mu.conv.weight' in keys and 'logvar.conv.bias' in keys
    x = torch.zeros(1, 3, 2, 5, 5)
    mu, logvar = enc(x)                      # plain torch pass-through: runs on the CPU
    assert mu.shape == (1, 8, 2, 5, 5) and logvar.shape == mu.shape


def test_zero_arena_and_per_sample_switch_host_logic():
    """host-side state of hpvg.ops that needs no GPU: the bump allocator behind zeros_small() and the per-draw BatchNorm switch"""
    import torch
    from hpvg import ops
    dev = torch.device("cpu")
    a = ops.zeros_small(10, dev)
    assert a.shape == (10,) and float(a.abs().sum()) == 0.0           # no arena: plain zeros
    with ops.zero_arena(dev, 128):
        x = ops.zeros_small(40, dev)
        y = ops.zeros_small(40, dev)
        assert x.untyped_storage().data_ptr() == y.untyped_storage().data_ptr()          # slices of one buffer
        assert y.data_ptr() - x.data_ptr() == 64 * 4                                      # 32-float granules
        z = ops.zeros_small(100, dev)                                                     # does not fit any more: falls back
        assert z.untyped_storage().data_ptr() != x.untyped_storage().data_ptr()
        x += 1.0
        assert float(y.sum()) == 0.0
    assert ops._ARENA[0] is None
    assert ops._BN_PER_SAMPLE[0] is False
    with ops.bn_per_sample(True):
        assert ops._BN_PER_SAMPLE[0] is True
        with ops.bn_per_sample(False):
            assert ops._BN_PER_SAMPLE[0] is False
        assert ops._BN_PER_SAMPLE[0] is True
    assert ops._BN_PER_SAMPLE[0] is False


def test_library_optimizer_refuses_cpu_parameters_and_keeps_torch_state_layout():
    """hpvg.optim.Adam has no CPU path; its param_groups carry torch.optim.Adam's keys (checkpoint compatibility)"""
    import torch
    from hpvg import optim
    from hpvg.lib import HpvgError
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    o = optim.Adam([{"params": [p], "lr": 1e-4}], lr=5e-4, betas=(0.5, 0.999))
    with pytest.raises(HpvgError):
        o.step()
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3))], lr=5e-4, betas=(0.5, 0.999)).state_dict()['param_groups'][0]
    mine = o.state_dict()['param_groups'][0]
    assert {'lr', 'betas', 'eps', 'weight_decay', 'amsgrad', 'maximize', 'params'} <= set(mine) <= set(ref)
    assert mine['lr'] == 1e-4 and mine['betas'] == (0.5, 0.999) and mine['eps'] == ref['eps']
    assert not optim.use_library_optimizer([p])


def test_library_optimizer_call_tables(monkeypatch):
    """host logic of hpvg.optim.Adam.step on CPU tensors with the library calls recorded instead of executed: tensors per call
    (<= 32), partial-sum slots, finalize / advance flags on the last chunk only, NULL moments for clipped-but-not-owned tensors,
    per-group learning rates, and the refusal of a changing gradient set"""
    import ctypes
    import torch
    from hpvg import lib, optim
    calls = []
    monkeypatch.setattr(optim.Adam, "_check", staticmethod(lambda t, what: None))
    monkeypatch.setattr(optim, "_stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(lib, "call", lambda name, *a: calls.append((name, a)))
    owned = [torch.nn.Parameter(torch.zeros(3 + i)) for i in range(40)]
    extra = [torch.nn.Parameter(torch.zeros(2)) for _ in range(5)]
    for p in owned + extra:
        p.grad = torch.ones_like(p)
    opt = optim.Adam([{"params": owned[:10], "lr": 1e-4}, {"params": owned[10:]}], lr=5e-4, betas=(0.5, 0.999))
    opt.step(clip_params=owned + extra, max_norm=5.0)
    clip = [a for n, a in calls if n == "hpvg_grad_clip_coef"]
    adam = [a for n, a in calls if n == "hpvg_adam_step"]
    assert [a[0] for a in clip] == [32, 13] and [a[0] for a in adam] == [32, 13]
    slots = 45 * lib.OPT_BLOCKS
    assert [(a[4], a[5], a[6]) for a in clip] == [(0, slots, 0), (32 * lib.OPT_BLOCKS, slots, 1)]      # slot_base, total, finalize
    assert [a[7] for a in clip] == [5.0, 5.0]
    assert [(a[10], a[11]) for a in adam] == [(1, 0), (1, 1)]                                          # use_clip, advance_step
    assert (adam[0][7], adam[0][8], adam[0][9]) == (0.5, 0.999, 1e-8)
    lrs = list(adam[0][6]) + list(adam[1][6])
    assert lrs[:10] == [pytest.approx(1e-4)] * 10 and lrs[10:40] == [pytest.approx(5e-4)] * 30 and lrs[40:] == [0.0] * 5
    numel = list(adam[0][5]) + list(adam[1][5])
    assert numel == [3 + i for i in range(40)] + [2] * 5
    moments = list(adam[1][3])
    assert all(m is not None for m in moments[:8]) and moments[8:] == [None] * 5                      # extra tensors: scaled only
    assert list(adam[1][1])[8:] == [None] * 5 and list(adam[1][2])[8:] == [e.grad.data_ptr() for e in extra]
    assert all(p._version == 1 for p in owned) and opt.last_launches == 4
    # no clipping: one kind of call, the extra tensors are not touched
    calls.clear()
    opt.step()
    assert [n for n, _ in calls] == ["hpvg_adam_step", "hpvg_adam_step"] and [a[0] for _, a in calls] == [32, 8]
    assert [(a[10], a[11]) for _, a in calls] == [(0, 0), (0, 1)]
    # one shared step count: a parameter that skips a step is refused
    owned[3].grad = None
    with pytest.raises(lib.HpvgError):
        opt.step()


def test_kernel_selection_rule():
    """hpvg_conv_forward's choice of kernel per layer (host logic): the brick tcgen05 kernel when its 4-slice units
    fill the SMs' unit slots and the depth is a multiple of 4, the column-streaming kernel otherwise; the 3-channel ends and
    odd channel counts never reach tcgen05 kernels they do not support"""
    from hpvg import lib
    prev = lib.set_conv_col_mode(-1)
    try:
        pick = lib.conv_kernel_choice
        assert pick(1, 64, 64, 16, 64, 64) == lib.KERNEL_TC_BRICK        # BASELINE configs[1], finest level: 128 full units
        assert pick(1, 64, 64, 32, 128, 128) == lib.KERNEL_TC_BRICK      # configs[4]: 1 024 units, 7 rounds at 0.99 fill
        assert pick(1, 64, 64, 13, 64, 64) == lib.KERNEL_TC_COLUMN       # default sampling rates: depth not a multiple of 4
        assert pick(1, 64, 64, 6, 54, 54) == lib.KERNEL_TC_COLUMN        # ragged bricks, half-empty second unit
        assert pick(1, 64, 64, 4, 32, 32) == lib.KERNEL_TC_COLUMN        # 16 units on 148 SMs
        assert pick(1, 128, 64, 4, 32, 32) == lib.KERNEL_TC_BRICK        # decoder head: the column kernel is 64-input-channel only
        assert pick(1, 64, 64, 1, 64, 64, kd=1) == lib.KERNEL_TC_BRICK   # 2-D layers (networks_2d): column kernel is 3-D only
        assert pick(1, 64, 3, 16, 64, 64, y_wide=False) == lib.KERNEL_TC_BRICK     # thin-output tails run the brick kernel's NOUT = 16 form
        assert pick(1, 3, 64, 16, 64, 64, x_wide=False) == lib.KERNEL_EXPAND       # head convs: TF32 mma.sync kernel
        assert pick(1, 8, 8, 4, 8, 8) == lib.KERNEL_DIRECT                         # the 8-channel test networks
        assert pick(1, 64, 64, 16, 64, 64, pad=3) < 0                              # invalid geometry is refused
        lib.set_conv_col_mode(0)
        assert pick(1, 64, 64, 6, 54, 54) == lib.KERNEL_TC_BRICK
        lib.set_conv_col_mode(1)
        assert pick(1, 64, 64, 16, 64, 64) == lib.KERNEL_TC_COLUMN
    finally:
        lib.set_conv_col_mode(prev)


def test_wgrad_workspace_plan():
    """hpvg_conv_wgrad_workspace (host logic): fp32 partials [splits][taps][Cin][Cout] with splits x KD x channel blocks <= 148
    CTAs, one split per group of bricks, plus 256 bytes for the arrival counter of the in-kernel reduction's grid barrier; large
    enough for either weight-gradient kernel form"""
    from hpvg import lib
    ws = lib.load().hpvg_conv_wgrad_workspace

    def splits(n, cin, cout, d, h, w, kd, pad):
        nbytes = ws(n, cin, cout, d, h, w, kd, pad, lib.FMT_NDHWC_BF16, lib.FMT_NDHWC_BF16) - 256    # + the grid barrier's counter slot
        per_split = kd * 9 * cin * cout * 4
        assert nbytes % per_split == 0
        return nbytes // per_split

    assert splits(1, 64, 64, 16, 64, 64, 3, 1) == 47       # 512 bricks, 148 // 3 = 49 wanted -> 11 bricks per split
    assert splits(1, 64, 64, 4, 32, 32, 3, 1) == 32        # 64 bricks, 2 per split
    assert splits(1, 128, 64, 4, 32, 32, 3, 1) == 16       # two input-channel blocks share the SMs
    assert splits(1, 64, 64, 1, 64, 64, 1, 1) == 32        # 2-D layer: 32 bricks, one each
    for shape in [(1, 64, 64, 16, 64, 64, 3, 1), (2, 64, 128, 7, 20, 33, 3, 1), (1, 64, 64, 22, 50, 50, 3, 0)]:
        s = splits(*shape)
        blocks = (shape[1] // 64) * (shape[2] // 64)
        assert 1 <= s and s * shape[6] * blocks <= 148


def test_peer_allreduce_argument_checks_need_no_gpu():
    """hpvg_peer_allreduce_avg validates ranks, bucket size and pointers on the host before it launches anything"""
    import ctypes
    from hpvg import lib
    handle = lib.load()
    two = (ctypes.c_void_p * 2)(None, None)
    assert handle.hpvg_peer_allreduce_avg(two, two, 0, lib.PEER_MAX_RANKS + 1, 64, None) == -1
    assert b"at most" in handle.hpvg_last_error()
    assert handle.hpvg_peer_allreduce_avg(two, two, 2, 2, 64, None) == -1                 # rank outside the world
    assert handle.hpvg_peer_allreduce_avg(two, two, 0, 2, 12, None) == -1                 # not a multiple of 4 x world floats
    assert b"multiple" in handle.hpvg_last_error()
    assert handle.hpvg_peer_allreduce_avg(two, two, 0, 2, 64, None) == -1                 # unmapped bucket
    assert b"not mapped" in handle.hpvg_last_error()
    assert handle.hpvg_peer_allreduce_avg(None, two, 0, 2, 64, None) == -1
    assert lib.PEER_HANDLE_BYTES == 64 and lib.PEER_SIGNAL_BYTES == 8192


def test_peer_bucket_layout_is_host_logic():
    """hpvg_peer_bucket_numel: every tensor starts on a float4 slot, the slots are dealt out evenly over the ranks"""
    from hpvg import lib
    handle = lib.load()
    assert handle.hpvg_peer_bucket_numel(3, lib.longlong_array([7, 5, 64]), 2) == 80          # 2 + 2 + 16 slots -> 10 per rank
    assert handle.hpvg_peer_bucket_numel(3, lib.longlong_array([7, 5, 64]), 8) == 96          # 20 slots -> 3 per rank
    assert handle.hpvg_peer_bucket_numel(1, lib.longlong_array([110592]), 8) == 110592
    assert handle.hpvg_peer_bucket_numel(1, lib.longlong_array([-1]), 2) == -1
    import ctypes
    two = (ctypes.c_void_p * 2)(None, None)
    assert handle.hpvg_peer_allreduce_avg_tensors(two, two, 0, 2, 64, 0, None, None, None) == -1
