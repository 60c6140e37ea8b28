"""Import the REFERENCE's model files in this container (build box only; needs /root/reference) by
stubbing the third-party packages they import but that are not installed here.
Used only by the golden-vector generators; nothing under tests/ that runs on the GPU box imports this."""
import importlib.util
import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference/VisionLLMv2/visionllmv2/model"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.util.spec_from_loader(name, loader=None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    class DropPath(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    def trunc_normal_(t, std=0.02, **k):
        return nn.init.trunc_normal_(t, std=std)

    def to_2tuple(x):
        return (x, x) if not isinstance(x, (tuple, list)) else tuple(x)

    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=DropPath, trunc_normal_=trunc_normal_, to_2tuple=to_2tuple)

    import transformers  # noqa: F401
    import transformers.pytorch_utils as pu
    import transformers.utils as tu
    for name, fn in (("find_pruneable_heads_and_indices", lambda *a, **k: None),
                     ("prune_linear_layer", lambda *a, **k: None),
                     ("apply_chunking_to_forward", lambda fn, cs, dim, *t: fn(*t)),
                     ("meshgrid", torch.meshgrid)):
        if not hasattr(pu, name):
            setattr(pu, name, fn)
    if not hasattr(tu, "is_ninja_available"):
        tu.is_ninja_available = lambda: False

    class Conv2d(nn.Conv2d):
        def __init__(self, *a, norm=None, activation=None, **k):
            super().__init__(*a, **k)
            self.norm, self.activation = norm, activation

        def forward(self, x):
            x = super().forward(x)
            if self.norm is not None:
                x = self.norm(x)
            if self.activation is not None:
                x = self.activation(x)
            return x

    class ShapeSpec:
        def __init__(self, channels=None, height=None, width=None, stride=None):
            self.channels, self.height, self.width, self.stride = channels, height, width, stride

    _mod("detectron2")
    _mod("detectron2.layers", Conv2d=Conv2d, ShapeSpec=ShapeSpec, get_norm=lambda n, c: nn.GroupNorm(32, c))
    _mod("fvcore")
    _mod("fvcore.nn")
    _mod("fvcore.nn.weight_init", c2_xavier_fill=lambda m: None)
    sys.modules["fvcore.nn"].weight_init = sys.modules["fvcore.nn.weight_init"]

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c

    _mod("mmcv")
    _mod("mmcv.runner", BaseModule=nn.Module, _load_checkpoint=lambda *a, **k: {}, load_checkpoint=lambda *a, **k: None)
    _mod("mmcv.cnn", build_norm_layer=lambda *a, **k: None, constant_init=lambda *a, **k: None,
         trunc_normal_init=lambda *a, **k: None)
    _mod("mmcv.cnn.bricks")
    _mod("mmcv.cnn.bricks.transformer", FFN=nn.Module, build_dropout=lambda *a, **k: nn.Identity())
    _mod("mmcv.cnn.utils")
    _mod("mmcv.cnn.utils.weight_init", trunc_normal_=trunc_normal_)
    _mod("mmcv.utils", to_2tuple=to_2tuple)
    _mod("mmdet")
    _mod("mmdet.utils", get_root_logger=lambda *a, **k: None)
    _mod("mmdet.models")
    _mod("mmdet.models.builder", BACKBONES=_Reg())
    _mod("transformers.models.deformable_detr.load_custom", load_cuda_kernels=lambda: None)


def load_file(pkg, name, path):
    full = f"{pkg}.{name}"
    spec = importlib.util.spec_from_file_location(full, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[full] = m
    spec.loader.exec_module(m)
    return m


def load_internvit():
    install_stubs()
    pkg = "refpkg_internvit"
    p = types.ModuleType(pkg); p.__path__ = [f"{REF}/internvit"]; sys.modules[pkg] = p
    cfg = load_file(pkg, "configuration_intern_vit", f"{REF}/internvit/configuration_intern_vit.py")
    try:
        load_file(pkg, "flash_attention", f"{REF}/internvit/flash_attention.py")
    except Exception:
        pass
    mod = load_file(pkg, "modeling_intern_vit", f"{REF}/internvit/modeling_intern_vit.py")
    return cfg, mod


def load_gdino():
    install_stubs()
    pkg = "refpkg_gdino"
    p = types.ModuleType(pkg); p.__path__ = [f"{REF}/grounding_dino"]; sys.modules[pkg] = p
    cfg = load_file(pkg, "configuration_grounding_dino", f"{REF}/grounding_dino/configuration_grounding_dino.py")
    mod = load_file(pkg, "modeling_ov_grounding_dino_mask_dn", f"{REF}/grounding_dino/modeling_ov_grounding_dino_mask_dn.py")
    return cfg, mod


def load_internlm2():
    install_stubs()
    pkg = "refpkg_internlm2"
    p = types.ModuleType(pkg); p.__path__ = [f"{REF}/internlm2"]; sys.modules[pkg] = p
    cfg = load_file(pkg, "configuration_internlm2", f"{REF}/internlm2/configuration_internlm2.py")
    mod = load_file(pkg, "modeling_internlm2", f"{REF}/internlm2/modeling_internlm2.py")
    return cfg, mod


def load_gdino_with_dcnv3():
    """`load_gdino()` plus the reference's DCNv3 module package bound to `gd.opsm` (gd.py:64-67 imports it relatively,
    which fails under the shim's flat package): lets `InternImage(core_op='DCNv3_pytorch')` -- the pure-PyTorch core,
    ops_dcnv3/modules/dcnv3.py:86-208 -- run on CPU.  The compiled `DCNv3` extension its functions file imports
    (dcnv3_func.py:16) is stubbed."""
    import logging
    cfgm, gd = load_gdino()
    sys.modules.setdefault("DCNv3", types.ModuleType("DCNv3"))
    base = f"{REF}/ops_dcnv3"
    pkg = types.ModuleType("refpkg_dcnv3"); pkg.__path__ = [base]; sys.modules["refpkg_dcnv3"] = pkg
    fpk = types.ModuleType("refpkg_dcnv3.functions"); fpk.__path__ = [base + "/functions"]
    sys.modules["refpkg_dcnv3.functions"] = fpk
    ff = load_file("refpkg_dcnv3.functions", "dcnv3_func", base + "/functions/dcnv3_func.py")
    fpk.DCNv3Function, fpk.dcnv3_core_pytorch = ff.DCNv3Function, ff.dcnv3_core_pytorch
    mpk = types.ModuleType("refpkg_dcnv3.modules"); mpk.__path__ = [base + "/modules"]
    sys.modules["refpkg_dcnv3.modules"] = mpk
    mm = load_file("refpkg_dcnv3.modules", "dcnv3", base + "/modules/dcnv3.py")
    mpk.DCNv3, mpk.DCNv3_pytorch = mm.DCNv3, mm.DCNv3_pytorch
    gd.opsm = mpk
    gd.get_root_logger = lambda *a, **k: logging.getLogger("ref")
    return cfgm, gd


def load_unipose():
    """The reference's UniPose model file as a package by path (relative imports of .utils / .ops resolve through the
    import system).  The compiled `MultiScaleDeformableAttention` extension its functions file imports
    (unipose/ops/functions/ms_deform_attn_func.py:19) is stubbed by a module whose forward is the reference's OWN
    pure-PyTorch core (`ms_deform_attn_core_pytorch`, same file :41-61), bound after the functions module loads."""
    import importlib
    install_stubs()
    ext = types.ModuleType("MultiScaleDeformableAttention")
    sys.modules["MultiScaleDeformableAttention"] = ext
    pkg = types.ModuleType("refpkg_unipose"); pkg.__path__ = [f"{REF}/unipose"]
    pkg.__spec__ = importlib.util.spec_from_loader("refpkg_unipose", loader=None, is_package=True)
    sys.modules["refpkg_unipose"] = pkg
    # gd-style relative import of ..ops_dcnv3 fails harmlessly (guarded by try/except in the file)
    fn = importlib.import_module("refpkg_unipose.ops.functions.ms_deform_attn_func")
    ext.ms_deform_attn_forward = lambda v, shapes, lsi, loc, w, step: fn.ms_deform_attn_core_pytorch(
        v, [(int(h), int(ww)) for h, ww in shapes.tolist()], loc, w)
    return importlib.import_module("refpkg_unipose.modeling_unipose")
