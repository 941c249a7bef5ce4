"""CPU-side checks: the C-ABI library loads and exports every symbol include/srcgan_b200.h declares,
the host logic (packed-weight caches, engine selection, drop-in modules, data-parallel helpers) works
without a GPU, and the product path refuses CPU tensors instead of falling back."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "srcgan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srcgan_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from srcgan_b200 import _lib
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.srcgan_version()


def test_conv_params_struct_matches_header():
    import ctypes as C
    from srcgan_b200._lib import ConvParams
    src = open(os.path.join(ROOT, "include", "srcgan_b200.h")).read()
    body = src[src.index("typedef struct srcgan_conv_params {"):src.index("} srcgan_conv_params;")]
    fields = re.findall(r"(?:int32_t|int64_t|float|const void\*|const float\*|void\*)\s+([a-z0-9_, ]+);", body)
    names = [n.strip() for grp in fields for n in grp.split(",")]
    assert names == [f[0] for f in ConvParams._fields_]
    assert C.sizeof(ConvParams) % 8 == 0


def test_integration_doc_struct_matches_binding():
    """The stand-alone ConvParams mirror printed in INTEGRATION.md is field for field (name and ctype) the one
    srcgan_b200/_lib.py binds - a reader who copies the document must not pass a short struct."""
    import ctypes as C
    from srcgan_b200._lib import ConvParams
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = doc[doc.index("class ConvParams(C.Structure)"):]
    body = body[:body.index("lib.srcgan_conv_fprop.argtypes")]
    fields = re.findall(r'\("([a-z0-9_]+)",\s*C\.(c_[a-z0-9_]+)\)', body)
    assert [n for n, _t in fields] == [n for n, _t in ConvParams._fields_]
    assert all(getattr(C, t) is bt for (_n, t), (_m, bt) in zip(fields, ConvParams._fields_))   # c_int32 is an alias of c_int

    class DocParams(C.Structure):
        _fields_ = [(n, getattr(C, t)) for n, t in fields]
    assert C.sizeof(DocParams) == C.sizeof(ConvParams) and C.sizeof(ConvParams) % 8 == 0


def test_argument_validation_without_gpu():
    """Bad arguments are rejected with an error code + message before any CUDA call."""
    import ctypes as C
    from srcgan_b200 import _lib
    lib = _lib.load()
    p = _lib.ConvParams()
    rc = lib.srcgan_conv_fprop(C.byref(p), None)
    assert rc == 1 and b"conv" in lib.srcgan_last_error()
    with pytest.raises(RuntimeError, match="non-positive"):
        _lib.check(rc, "conv_fprop")
    assert lib.srcgan_pack_weights(None, 1, 1, 3, 3, 0, 0, None, None) == 1


def test_introspection_entry_points_without_gpu():
    """launch counter and last-kernel name are callable before any launch (bench.py keys its roofline record on them)."""
    from srcgan_b200 import _lib
    assert _lib.launch_count() >= 0
    assert isinstance(_lib.last_kernel(), str)


def test_thin_layers_are_padded_onto_the_tensor_core_engine():
    """engine.select() sends the 1- and 3-channel sides of the discriminator to SIMT; nn.py pads them to 16/32/64
    channels for the tcgen05 kernels - the padded shapes must be ones the tensor-core engine accepts."""
    import torch
    from srcgan_b200 import engine
    from srcgan_b200._lib import ENGINE_SIMT, ENGINE_TC
    bf = torch.bfloat16
    assert engine.select(256, 1, 4, 1, False, bf, 62, 62)[0] == ENGINE_SIMT
    assert engine.select(256, 32, 4, 1, False, bf, 62, 62)[0] == ENGINE_TC           # patch logits, cout padded to 32
    assert engine.select_wgrad(256, 1, 4, 1, False, bf, 62, 62) == ENGINE_SIMT
    assert engine.select_wgrad(256, 64, 4, 1, False, bf, 62, 62) == ENGINE_TC        # its wgrad, dY padded to 64
    assert engine.select_wgrad(3, 64, 4, 2, False, bf, 128, 128) == ENGINE_SIMT
    assert engine.select_wgrad(16, 64, 4, 2, False, bf, 128, 128) == ENGINE_TC       # first layer wgrad, image padded to 16
    assert not engine.tc_dgrad_s2_supported(3, 64, 4, 2, 1, bf)
    assert engine.tc_dgrad_s2_supported(32, 64, 4, 2, 1, bf)                         # its dgrad, image padded to 32


def test_thin_input_layers_take_the_tiled_ffma_kernel(monkeypatch):
    """<= 4 input channels through a 4x4 filter (first discriminator layer, gradient of the patch-logit layer) go to the FFMA
    kernel thin_in_tiled; 3x3 stride-1 image convolutions stay on the tcgen05 sweep (K zero-filled by TMA)."""
    import torch
    from srcgan_b200 import engine
    from srcgan_b200._lib import ENGINE_SIMT, ENGINE_TC, WL_RSCK
    bf = torch.bfloat16
    assert engine.select(3, 64, 4, 2, False, bf, 128, 128) == (ENGINE_SIMT, WL_RSCK)
    assert engine.select(1, 256, 4, 1, False, bf, 62, 62) == (ENGINE_SIMT, WL_RSCK)
    assert engine.select(3, 64, 3, 1, False, bf, 256, 256)[0] == ENGINE_TC
    assert engine.select(3, 32, 4, 2, False, bf, 128, 128)[0] == ENGINE_TC           # not a multiple of 64 output channels
    monkeypatch.setenv("SRCGAN_B200_NO_THIN_TILED", "1")
    assert engine.select(3, 64, 4, 2, False, bf, 128, 128)[0] == ENGINE_TC


def test_operator_layer_thin_tensors_are_pitch_8_views(monkeypatch):
    """functional.py: a bf16 tensor with fewer than 8 channels is the view buf[..., :c] of a pitch-8 buffer (TMA-addressable),
    recognised by _sl() / _canon(); fp32 tensors and wide tensors stay dense; .contiguous() gives an ordinary dense copy."""
    import torch
    from srcgan_b200 import functional as Fn
    like = torch.empty((), dtype=torch.bfloat16)
    t = Fn._new_thin(2, 5, 7, 3, like)
    assert tuple(t.shape) == (2, 5, 7, 3) and t.stride() == (280, 56, 8, 1) and not t.is_contiguous()
    assert Fn._is_padded_view(t) and Fn._canon(t) is t
    s = Fn._sl(t)
    assert (s.ld, s.c0, s.c, s.ptr) == (8, 0, 3, t.data_ptr()) and tuple(s.buf.shape) == (2, 5, 7, 8)
    t.copy_(torch.arange(2 * 5 * 7 * 3, dtype=torch.float32).reshape(2, 5, 7, 3) % 17)
    assert torch.equal(s.buf[..., :3], t) and t.contiguous().is_contiguous()
    assert Fn._new_thin(2, 5, 7, 3, torch.empty((), dtype=torch.float32)).is_contiguous()      # fp32 parity mode: dense
    assert Fn._new_thin(2, 5, 7, 8, like).is_contiguous()                                      # 8 channels: dense
    odd = torch.empty(2, 5, 7, 6, dtype=torch.bfloat16)[..., :3]                               # some other strided view
    assert not Fn._is_padded_view(odd) and Fn._canon(odd).is_contiguous()
    wide = torch.empty(2, 5, 7, 64, dtype=torch.bfloat16)
    assert Fn._sl(wide).ld == 64 and Fn._canon(wide) is wide
    monkeypatch.setenv("SRCGAN_B200_NO_THIN_PITCH", "1")
    assert Fn._new_thin(2, 5, 7, 3, like).is_contiguous()


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from srcgan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_networks_refuse_cpu_tensors():
    from srcgan_b200 import losses, nn as snn
    for net, shape in ((snn.NLayerDiscriminator(3, 64, 2), (1, 3, 32, 32)),
                       (snn.RDDBNetB(3, 3, 64, nb=1, mode="x4"), (1, 3, 8, 8))):
        with pytest.raises(RuntimeError, match="no CPU"):
            net(torch.rand(*shape))
    with pytest.raises(RuntimeError, match="no CPU"):
        losses.L1Loss()(torch.rand(4), torch.rand(4))


def test_state_dict_compatibility_with_oracle_layout():
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    g = snn.RDDBNetB(3, 3, 64, nb=3, mode="x4")
    assert list(g.state_dict().keys()) == list(O.init_rddbnet_b(0).keys())
    g.load_state_dict(O.init_rddbnet_b(3), strict=True)
    a = snn.RDDBNetA(3, 3, 64, nb=3, mode="x4")
    a.load_state_dict(O.init_rddbnet_a(3), strict=True)
    d = snn.NLayerDiscriminator(3, 64, 2)
    d.load_state_dict(O.init_discriminator(3), strict=True)
    assert sum(p.numel() for p in g.parameters()) == 2309507
    assert sum(p.numel() for p in a.parameters()) == 3195715
    assert sum(p.numel() for p in d.parameters()) == 663361


def test_dense_block_transposed_weights_are_the_adjoint():
    """Host logic of the mirrored dense block: step-k virtual weights reproduce autograd's input
    gradient of the reference dense block (checked with torch on the CPU)."""
    import torch.nn.functional as F
    from oracle import srcgan_oracle as O
    from srcgan_b200.nn import _wT
    torch.manual_seed(0)
    nf, gc = 16, 8
    ws = [torch.randn(gc if k < 5 else nf, nf + gc * (k - 1), 3, 3) * 0.1 for k in range(1, 6)]
    x = torch.randn(1, nf, 6, 5, requires_grad=True)
    feats = [x]
    for k in range(4):
        feats.append(F.leaky_relu(F.conv2d(torch.cat(feats, 1), ws[k], padding=1), 0.2))
    out = F.conv2d(torch.cat(feats, 1), ws[4], padding=1) * 0.2 + x
    g = torch.randn_like(out)
    out.backward(g)
    # mirrored pass
    D = [g]                                     # [dOut, dZ4, dZ3, dZ2, dZ1]
    for k in (4, 3, 2, 1):
        sl = slice(nf + gc * (k - 1), nf + gc * k)
        blocks = [_wT(ws[4][:, sl]) * 0.2] + [_wT(ws[j - 1][:, sl]) for j in range(4, k, -1)]
        d = F.conv2d(torch.cat(D, 1), torch.cat(blocks, 1), padding=1)
        mask = torch.where(feats[k] > 0, torch.ones_like(d), torch.full_like(d, 0.2))
        D.append(d * mask)
    blocks = [_wT(ws[4][:, :nf]) * 0.2] + [_wT(ws[j - 1][:, :nf]) for j in range(4, 0, -1)]
    dx = F.conv2d(torch.cat(D, 1), torch.cat(blocks, 1), padding=1) + g
    assert torch.allclose(dx, x.grad, rtol=1e-4, atol=1e-5)


def test_engine_selection():
    from srcgan_b200 import engine
    from srcgan_b200._lib import ENGINE_SIMT, ENGINE_TC, WL_RSCK, WL_TC
    bf, f32 = torch.bfloat16, torch.float32
    assert engine.select(64, 32, 3, 1, False, bf, 64, 64) == (ENGINE_TC, WL_TC)
    assert engine.select(64, 32, 3, 1, False, f32, 64, 64) == (ENGINE_SIMT, WL_RSCK)      # fp32 parity mode
    assert engine.select(3, 64, 3, 1, False, bf, 64, 64) == (ENGINE_TC, WL_TC)            # thin image convs: K / N padded
    assert engine.select(64, 3, 3, 1, False, bf, 64, 64)[0] == ENGINE_TC
    assert engine.select(256, 1, 4, 1, False, bf, 62, 62)[0] == ENGINE_SIMT               # 4x4 patch-logit conv
    assert engine.select(3, 64, 3, 1, False, f32, 64, 64)[0] == ENGINE_SIMT
    assert engine.select_wgrad(192, 64, 3, 1, False, bf, 64, 64) == ENGINE_TC
    assert engine.select_wgrad(3, 64, 3, 1, False, bf, 64, 64) == ENGINE_TC
    assert engine.select_wgrad(3, 64, 4, 2, False, bf, 64, 64) == ENGINE_SIMT


def test_dropin_modules_resolve_reference_imports():
    """`from model import RDDBNetA, RDDBNetB, NLayerDiscriminator, ...` (train.py:11), `import losses`,
    `import metrics` resolve to this package when srcgan_b200/dropin is first on the path."""
    code = (
        "import sys; sys.path.insert(0, %r);"
        "from model import RDDBNetA, RDDBNetB, NLayerDiscriminator, SRDenseNetA, SRDenseNetB;"
        "import losses, metrics;"
        "assert RDDBNetB.__module__ == 'srcgan_b200.nn';"
        "assert repr(losses.L1Loss()) == 'L1' and repr(metrics.SSIM()) == 'SSIM' and repr(metrics.AE()) == 'AE';"
        "assert repr(losses.PSNRLoss()) == 'PSNR' and repr(losses.MSELoss()) == 'MSE';"
        "print('ok')" % os.path.join(ROOT, "srcgan_b200", "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not on this box")
def test_reference_train_script_constructs_on_the_dropin():
    """The reference's UNMODIFIED train.py imports and builds SRCycleGAN on the drop-in modules
    (visdom / skimage / dataset are stubbed: they are I/O, outside the hot path)."""
    code = (
        "import sys, types, torch;"
        "sys.path.insert(0, '/root/reference/src'); sys.path.insert(0, %r);"
        "sys.modules['visdom'] = types.SimpleNamespace(Visdom=lambda *a, **k: None);"
        "sk = types.ModuleType('skimage'); sk.io = types.ModuleType('skimage.io'); sk.color = types.ModuleType('skimage.color');"
        "sk.color.lab2rgb = sk.color.rgb2lab = sk.color.rgb2gray = None; sk.io.imsave = sk.io.imread = None;"
        "sys.modules.update({'skimage': sk, 'skimage.io': sk.io, 'skimage.color': sk.color});"
        "import train;"
        "opt = train.params(); opt.device = torch.device('cpu'); opt.mode = 'x4';"
        "m = train.SRCycleGAN(opt);"
        "assert type(m.netG_A).__module__ == 'srcgan_b200.nn' and type(m.netD_A).__module__ == 'srcgan_b200.nn';"
        "assert type(m.criterionCycle).__module__ == 'srcgan_b200.losses';"
        "assert len(m.optimizer_G.param_groups[0]['params']) == 102 + 118;"
        "print('ok')" % os.path.join(ROOT, "srcgan_b200", "dropin"))
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-3000:]


def test_checkpoint_names_follow_the_reference_scripts():
    from srcgan_b200 import checkpoint as C
    assert C.cas_checkpoint_name("RDDBNet", "A2C", 4, 7) == "RDDBNet_A2C_x4_0007.pth"            # trainCas.py:222
    assert C.cas_checkpoint_name("ResDeconv", "C2B", 2, 50, lab=True) == "ResDeconv@G2LAB_C2B_x2_0050.pth"
    n = C.parse_cas_checkpoint_name("./checkpoints/ResDeconv@G2LAB_C2B_x2_0050.pth")
    assert n == C.CasName("ResDeconv", True, "C2B", 2, 50)
    # the reference's own parse (testCas.py:40-41,52) agrees
    parts = "RDDBNet_A2C_x4_0007.pth".split(".pth")[0].split("_")
    assert (parts[0], int(parts[2][1])) == ("RDDBNet", C.parse_cas_checkpoint_name("RDDBNet_A2C_x4_0007.pth").up)
    assert C.cyclegan_checkpoint_names("x4", 3) == ("netG_A2B_SRtask_x4_0003.pth", "netG_B2A_SRtask_x4_0003.pth")
    with pytest.raises(ValueError):
        C.parse_cas_checkpoint_name("netG_A2B_SRtask_x4_0003.pth")


def _zoo_pairs():
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn, zoo
    return {
        "ResDeconv": (lambda: zoo.ResDeconv(1, 3), O.init_resdeconv(1, 3), 14982912),
        "EDSR": (lambda: zoo.EDSR(1, 1, 2, num_residuals=3), O.init_edsr(2, 1, 1, 2, num_residuals=3), None),
        "SRDenseNetA": (lambda: zoo.SRDenseNetA(1, 3, mode="x4", num_blocks=2, num_layers=2),
                        O.init_srdensenet(3, "A", 1, 3), None),
        "SRDenseNetB": (lambda: zoo.SRDenseNetB(3, 1, mode="x4", num_blocks=2, num_layers=2),
                        O.init_srdensenet(4, "B", 3, 1), None),
        "RDDBNet": (lambda: snn.RDDBNet(1, 1, 4), O.init_rddbnet_pkg(5, 1, 1, 4), 2229184),
        "SRDN": (lambda: snn.SRDN(1, 1, 2), O.init_srdn(6), None),
        "ESPCN": (lambda: snn.ESPCN(1, 1, 2), O.init_espcn(7, 1, 1, 2), None),
        "SRCNN": (lambda: snn.SRCNN(1, 3, 2), O.init_srcnn(8, 1, 3), None),
    }


@pytest.mark.parametrize("name", ["ResDeconv", "EDSR", "SRDenseNetA", "SRDenseNetB", "RDDBNet", "SRDN", "ESPCN", "SRCNN"])
def test_checkpoint_round_trip(tmp_path, name):
    """.pth files written here hold reference-keyed CPU tensors and load back strictly; where the reference tree
    is present, the REAL reference module loads the same file with strict=True (interchange both ways)."""
    import torch
    from srcgan_b200 import checkpoint as C
    ctor, sd, nparams = _zoo_pairs()[name]
    net = ctor()
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.load_state_dict(sd, strict=True)
    if nparams is not None:
        assert sum(p.numel() for p in net.parameters()) == nparams
    path = str(tmp_path / "checkpoints" / C.cas_checkpoint_name(name, "A2C", 2, 1))
    C.save_state(net, path)
    back = torch.load(path)
    assert all(torch.equal(back[k], sd[k]) for k in sd) and list(back) == list(sd)
    other = ctor()
    C.load_state(other, path)
    assert all(torch.equal(a, b) for a, b in zip(other.state_dict().values(), sd.values()))
    if os.path.isdir("/root/reference/src"):
        from oracle import ref_harness
        pkg, M, _l, _m = ref_harness.import_reference()
        ref = {"ResDeconv": lambda: pkg.ResDeconv(1, 3), "EDSR": lambda: pkg.EDSR(1, 1, 2, num_residuals=3),
               "SRDenseNetA": lambda: M.SRDenseNetA(1, 3, mode="x4", num_blocks=2, num_layers=2),
               "SRDenseNetB": lambda: M.SRDenseNetB(3, 1, mode="x4", num_blocks=2, num_layers=2),
               "RDDBNet": lambda: pkg.RDDBNet(1, 1, 4), "SRDN": lambda: pkg.SRDN(1, 1, 2),
               "ESPCN": lambda: pkg.ESPCN(1, 1, 2), "SRCNN": lambda: pkg.SRCNN(1, 3, 2)}[name]()
        ref.load_state_dict(torch.load(path), strict=True)
        assert list(ref.state_dict().keys()) == list(net.state_dict().keys())


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not on this box")
def test_reference_traincas_script_constructs_on_the_dropin():
    """The reference's UNMODIFIED trainCas.py (`from model import *`, eval(opt.SRModel), default CModel ResDeconv)
    builds CasSRC on the drop-in package for every generator name of src/model/__init__.py."""
    code = (
        "import sys, types, torch;"
        "sys.path.insert(0, '/root/reference/src'); sys.path.insert(0, %r);"
        "sys.modules['visdom'] = types.SimpleNamespace(Visdom=lambda *a, **k: None);"
        "sk = types.ModuleType('skimage'); sk.io = types.ModuleType('skimage.io'); sk.color = types.ModuleType('skimage.color');"
        "sk.color.lab2rgb = sk.color.rgb2lab = sk.color.rgb2gray = None; sk.io.imsave = sk.io.imread = None;"
        "sys.modules.update({'skimage': sk, 'skimage.io': sk.io, 'skimage.color': sk.color});"
        "import trainCas;"
        "opt = trainCas.params(); opt.device = torch.device('cpu'); opt.up = 2;"
        "seen = set();\n"
        "for sr in ('ESPCN', 'SRCNN', 'EDSR', 'RDDBNet', 'SRDN'):\n"
        "    opt.SRModel, opt.CModel = sr, 'ResDeconv'\n"
        "    m = trainCas.CasSRC(opt)\n"
        "    seen.add(type(m.netG_A2C).__module__); seen.add(type(m.netG_C2B).__module__)\n"
        "assert seen == {'srcgan_b200.nn', 'srcgan_b200.zoo'}, seen\n"
        "print('ok')" % os.path.join(ROOT, "srcgan_b200", "dropin"))
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-3000:]


def test_update_lr_mirrors_the_reference_epoch_schedule():
    """train.py:196-213 / trainCas.py:45-61: a fresh scheduler per call, stepped once - with the default 'cosine' policy the
    learning rate is multiplied by (1 + cos(pi / num_epochs)) / 2 per epoch.  Compared with the closed form and, where the
    reference tree is present, with the reference's own ``update_lr`` on identical optimizers."""
    import math
    import torch
    from srcgan_b200 import trainer
    opt = trainer.params()
    lin = torch.nn.Linear(2, 2)
    opts = [torch.optim.Adam(lin.parameters(), lr=1e-4, betas=(0.5, 0.999)), torch.optim.Adam(lin.parameters(), lr=1e-5)]
    for _ in range(3):
        trainer.update_lr(opts, opt)
    f = (1 + math.cos(math.pi / opt.num_epochs)) / 2
    assert opts[0].param_groups[0]["lr"] == pytest.approx(1e-4 * f ** 3, rel=1e-9)
    assert opts[1].param_groups[0]["lr"] == pytest.approx(1e-5 * f ** 3, rel=1e-9)
    opt.lr_policy = "linear"
    with pytest.raises(NotImplementedError):
        trainer.update_lr(opts, opt)
    assert hasattr(trainer.SRCycleGAN, "update_lr")
    from srcgan_b200 import trainer_cas
    assert hasattr(trainer_cas.CasSRC, "update_lr")
    if os.path.isdir("/root/reference/src"):
        from oracle import ref_harness
        train = ref_harness.import_train()

        class Holder:                                  # update_lr only touches self.optimizers
            pass
        h = Holder()
        h.optimizers = [torch.optim.Adam(lin.parameters(), lr=1e-4, betas=(0.5, 0.999)), torch.optim.Adam(lin.parameters(), lr=1e-5)]
        ropt = train.params()
        for _ in range(3):
            train.SRCycleGAN.update_lr(h, ropt)
        assert h.optimizers[0].param_groups[0]["lr"] == pytest.approx(1e-4 * f ** 3, rel=1e-9)
        assert h.optimizers[1].param_groups[0]["lr"] == pytest.approx(1e-5 * f ** 3, rel=1e-9)


def test_device_prefetcher_refuses_cpu():
    from srcgan_b200 import data
    with pytest.raises(RuntimeError, match="CUDA device"):
        data.DevicePrefetcher([torch.zeros(2)], "cpu")
