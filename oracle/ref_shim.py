"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules from /root/reference.

This file never ships with the product path.  It exists so that, in the build container (where
``/root/reference`` is mounted read-only), the CPU restatement in ``oracle/decoder_ref.py`` can be
pinned against the reference's own PyTorch code and so that ``oracle/make_golden.py`` can generate the
committed fixtures under ``tests/golden/``.  On the GPU box ``/root/reference`` does not exist and
``available()`` returns False; nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.

The reference decoder files import four names from packages that are not installed here
(``detectron2.config.configurable``, ``detectron2.layers.Conv2d``, ``detectron2.utils.registry.Registry``,
``fvcore.nn.weight_init.c2_xavier_fill``; video_mask2former_transformer_decoder.py:4,10-12).  They are
stubbed in ``sys.modules``; the reference sources are loaded by path and are
never copied into this repository.
"""
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("OPENVIS_REFERENCE", "/root/reference")
_DEC_DIR = os.path.join(REF_ROOT, "openvis/modeling/transformer_decoder")
_CLIP_DIR = os.path.join(REF_ROOT, "openvis/modeling/clip_adapter")
_MACLIP = os.path.join(REF_ROOT, "third_parties/mask_adapted_clip/mask_adapted_clip/model.py")

_loaded = {}


def available() -> bool:
    return os.path.isfile(os.path.join(_DEC_DIR, "video_mask2former_transformer_decoder.py"))


def _mod(name, **kw):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    m.__dict__.update(kw)
    return m


class _Registry(dict):
    """Stand-in for detectron2.utils.registry.Registry (dict with a .register decorator)."""

    def __init__(self, name):
        super().__init__()
        self.name = name

    def register(self, obj=None):
        def deco(o):
            self[o.__name__] = o
            return o

        return deco if obj is None else deco(obj)


class _Conv2d(torch.nn.Conv2d):
    """Stand-in for detectron2.layers.Conv2d, restated from its published definition (detectron2/layers/wrappers.py):
    torch.nn.Conv2d with optional `norm` and `activation` applied after the convolution."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = torch.nn.functional.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class _ShapeSpec:
    """Stand-in for detectron2.layers.ShapeSpec (channels / height / width / stride record)."""

    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


def _get_norm(norm, out_channels):
    """detectron2.layers.get_norm for the two values the shipped configs use: "" -> None, "GN" -> GroupNorm(32, C)."""
    if norm is None or norm == "":
        return None
    if norm == "GN":
        return torch.nn.GroupNorm(32, out_channels)
    raise NotImplementedError(norm)


def _c2_xavier_fill(m):
    torch.nn.init.kaiming_uniform_(m.weight, a=1)
    if m.bias is not None:
        torch.nn.init.constant_(m.bias, 0)


def _install_stubs():
    if "detectron2" in sys.modules and not getattr(sys.modules["detectron2"], "_ovis_stub", False):
        return  # a real detectron2 is importable: leave it alone
    _mod("detectron2", _ovis_stub=True)
    _mod("detectron2.config", configurable=lambda f=None, **k: f)
    _mod("detectron2.layers", Conv2d=_Conv2d, ShapeSpec=_ShapeSpec, get_norm=_get_norm)
    _mod("detectron2.utils")
    _mod("detectron2.utils.registry", Registry=_Registry)
    _mod("detectron2.utils.comm", get_local_rank=lambda: 0, synchronize=lambda: None)
    _mod("fvcore")
    _mod("fvcore.nn")
    _mod("fvcore.nn.weight_init", c2_xavier_fill=_c2_xavier_fill)


def _load(pkg_name, pkg_dir, name):
    full = f"{pkg_name}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(pkg_dir, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[full] = m
    spec.loader.exec_module(m)
    return m


def decoders():
    """Returns a namespace with the reference decoder classes (unmodified)."""
    if "dec" in _loaded:
        return _loaded["dec"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    pkg = _mod("refdec")
    pkg.__path__ = [_DEC_DIR]
    pe = _load("refdec", _DEC_DIR, "position_encoding")
    v = _load("refdec", _DEC_DIR, "video_mask2former_transformer_decoder")
    f = _load("refdec", _DEC_DIR, "frame_mask2former_transformer_decoder")
    sf = _load("refdec", _DEC_DIR, "side_adapter_frame_mask2former_transformer_decoder")
    sv = _load("refdec", _DEC_DIR, "side_adapter_video_mask2former_transformer_decoder")
    ns = types.SimpleNamespace(
        position_encoding=pe,
        video=v,
        frame=f,
        san_frame=sf,
        san_video=sv,
        VideoMultiScaleMaskedTransformerDecoder=v.VideoMultiScaleMaskedTransformerDecoder,
        FrameMultiScaleMaskedTransformerDecoder=f.FrameMultiScaleMaskedTransformerDecoder,
        SideAdapterFrameMultiScaleMaskedTransformerDecoder=sf.SideAdapterFrameMultiScaleMaskedTransformerDecoder,
        SideAdapterVideoMultiScaleMaskedTransformerDecoder=sv.SideAdapterVideoMultiScaleMaskedTransformerDecoder,
    )
    _loaded["dec"] = ns
    return ns


def decoder_kwargs(num_queries=100, hidden_dim=256, nheads=8, dim_feedforward=2048, dec_layers=9,
                   num_classes=1, mask_dim=None, num_frames=2):
    return dict(in_channels=hidden_dim, mask_classification=True, num_classes=num_classes,
                hidden_dim=hidden_dim, num_queries=num_queries, nheads=nheads,
                dim_feedforward=dim_feedforward, dec_layers=dec_layers, pre_norm=False,
                mask_dim=hidden_dim if mask_dim is None else mask_dim,
                enforce_input_project=False, num_frames=num_frames)


def side_adapter_module():
    """Reference ``SideAdapter`` with the vendored mask_adapted_clip model standing in for OpenAI clip
    (SURVEY.md Appendix A).  ``build_clip_model`` is replaced by a seeded random-init ViT-B/16 CLIP."""
    if "san" in _loaded:
        return _loaded["san"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    spec = importlib.util.spec_from_file_location("mask_adapted_clip.model", _MACLIP)
    clip_model = importlib.util.module_from_spec(spec)
    pkg = _mod("mask_adapted_clip")
    pkg.__path__ = [os.path.dirname(_MACLIP)]
    sys.modules["mask_adapted_clip.model"] = clip_model
    spec.loader.exec_module(clip_model)
    pkg.model = clip_model

    def _no_tokenizer(*a, **k):
        raise RuntimeError("clip.tokenize unavailable offline (ftfy missing); pass a text matrix")

    _mod("clip", model=clip_model, tokenize=_no_tokenizer)
    sys.modules["clip.model"] = clip_model
    cpkg = _mod("refclip")
    cpkg.__path__ = [_CLIP_DIR]
    utils = _load("refclip", _CLIP_DIR, "utils")
    sa = _load("refclip", _CLIP_DIR, "side_adapter")

    def build_clip_model(*_a, **_k):
        g = torch.random.get_rng_state()
        torch.manual_seed(5)
        m = clip_model.CLIP(512, 224, 12, 768, 16, 0, 77, 49408, 512, 8, 12).float().eval()
        torch.random.set_rng_state(g)
        return m

    sa.build_clip_model = build_clip_model
    utils.build_clip_model = build_clip_model
    ns = types.SimpleNamespace(module=sa, SideAdapter=sa.SideAdapter, clip_model=clip_model)
    _loaded["san"] = ns
    return ns


class _BitMasks:
    """Stand-in for detectron2.structures.BitMasks (not installed here), restated from its published definition
    (detectron2/structures/masks.py:88-209): `.tensor` [M, H, W] bool; get_bounding_boxes() -> object with `.tensor` [M, 4]
    = [x_min, y_min, x_max + 1, y_max + 1] of the non-zero pixels, zeros for an empty mask."""

    def __init__(self, tensor):
        self.tensor = torch.as_tensor(tensor).to(torch.bool)

    def get_bounding_boxes(self):
        boxes = torch.zeros(self.tensor.shape[0], 4, dtype=torch.float32)
        x_any = torch.any(self.tensor, dim=1)
        y_any = torch.any(self.tensor, dim=2)
        for idx in range(self.tensor.shape[0]):
            x = torch.where(x_any[idx, :])[0]
            y = torch.where(y_any[idx, :])[0]
            if len(x) > 0 and len(y) > 0:
                boxes[idx, :] = torch.as_tensor([x[0], y[0], x[-1] + 1, y[-1] + 1], dtype=torch.float32)
        return types.SimpleNamespace(tensor=boxes)


def clip_adapter(visual_state_dict=None):
    """The reference's ClipAdapter (openvis/modeling/clip_adapter/adapter.py:34-147), unmodified, around the vendored
    mask_adapted_clip CLIP (ViT-B/16) -- `visual_state_dict` (e.g. synthetic.seeded_clip_visual_params) replaces the visual
    tower's random initialisation.  torchvision's roi_align is the real one; detectron2's BitMasks is `_BitMasks` above.
    Returns the adapter instance (CPU, eval)."""
    san = side_adapter_module()                   # installs the clip / refclip packages and the seeded build_clip_model
    _mod("detectron2.structures", BitMasks=_BitMasks)
    ad = _load("refclip", _CLIP_DIR, "adapter")
    ad.build_clip_model = sys.modules["refclip.utils"].build_clip_model
    a = ad.ClipAdapter("ViT-B/16").eval()
    if visual_state_dict is not None:
        missing, unexpected = a.clip_model.visual.load_state_dict(visual_state_dict, strict=False)
        assert not unexpected and set(missing) <= {"mask_embedding"}, (missing, unexpected)
    return a


class on_cpu:
    """Context for running reference code that hard-codes `.cuda()` (adapter.py:92) in this GPU-less container, and --
    `fp32=True` -- for evaluating its `.half()` sections (adapter.py:106, 109) in fp32."""

    def __init__(self, fp32=False):
        self.fp32 = fp32

    def __enter__(self):
        self._cuda, self._half = torch.Tensor.cuda, torch.Tensor.half
        torch.Tensor.cuda = lambda t, *a, **k: t
        if self.fp32:
            torch.Tensor.half = lambda t, *a, **k: t.float()
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda, torch.Tensor.half = self._cuda, self._half
        return False


def msda_core_pytorch():
    """The reference's own pure-PyTorch multi-scale deformable attention (``ms_deform_attn_core_pytorch``,
    ops/functions/ms_deform_attn_func.py:55-77).  Its module refuses to import without the compiled extension, so an
    empty stand-in for ``MultiScaleDeformableAttention`` is registered first (the debug function never touches it)."""
    if "msda" in _loaded:
        return _loaded["msda"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _mod("MultiScaleDeformableAttention")
    d = os.path.join(REF_ROOT, "openvis/modeling/pixel_decoder/ops/functions")
    pkg = _mod("refmsda")
    pkg.__path__ = [d]
    m = _load("refmsda", d, "ms_deform_attn_func")
    _loaded["msda"] = m.ms_deform_attn_core_pytorch
    return _loaded["msda"]


def msda_module():
    """The reference's MSDeformAttn module class (ops/modules/ms_deform_attn.py:35-125), unmodified.  Without the compiled
    extension its forward takes its own `except:` branch (:118-121), i.e. ms_deform_attn_core_pytorch."""
    if "msda_module" in _loaded:
        return _loaded["msda_module"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _mod("MultiScaleDeformableAttention")
    ops = os.path.join(REF_ROOT, "openvis/modeling/pixel_decoder/ops")
    pkg = _mod("refops")
    pkg.__path__ = [ops]
    m = importlib.import_module("refops.modules.ms_deform_attn")
    _loaded["msda_module"] = m.MSDeformAttn
    return m.MSDeformAttn


def pixel_decoder():
    """The reference's MSDeformAttnPixelDecoder / MSDeformAttnTransformerEncoderOnly (openvis/modeling/pixel_decoder/
    msdeformattn.py:38-380), unmodified; its MSDeformAttn runs ms_deform_attn_core_pytorch (no compiled extension)."""
    if "pixdec" in _loaded:
        return _loaded["pixdec"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    _mod("MultiScaleDeformableAttention")
    m = sys.modules.get("detectron2.modeling")
    if m is None or not hasattr(m, "SEM_SEG_HEADS_REGISTRY"):
        _mod("detectron2.modeling", SEM_SEG_HEADS_REGISTRY=_Registry("SEM_SEG_HEADS"))
    d = os.path.join(REF_ROOT, "openvis/modeling/pixel_decoder")
    _mod("refpix").__path__ = [d]
    _mod("refpix.ops").__path__ = [os.path.join(d, "ops")]
    mod = importlib.import_module("refpix.msdeformattn")
    ns = types.SimpleNamespace(MSDeformAttnPixelDecoder=mod.MSDeformAttnPixelDecoder,
                               MSDeformAttnTransformerEncoderOnly=mod.MSDeformAttnTransformerEncoderOnly,
                               ShapeSpec=_ShapeSpec)
    _loaded["pixdec"] = ns
    return ns


def temporal():
    """The reference's temporal-association code, unmodified: ``match_via_embeds`` /
    ``batch_video_match_via_embeds`` (openvis/modeling/minvis.py:28-72), ``batch_index`` (openvis/utils/index.py:4-19)
    and ``TemporalInstanceResampler`` (openvis/modeling/resampler.py:189-323).  ``minvis.py`` also defines the MinVIS
    meta-architecture, whose base class and Detectron2 imports are stubbed (never instantiated here)."""
    if "temporal" in _loaded:
        return _loaded["temporal"]
    dec = decoders()
    _mod("detectron2.modeling", META_ARCH_REGISTRY=_Registry("META_ARCH"))
    _mod("detectron2.modeling.backbone", Backbone=object)
    _mod("detectron2.structures", ImageList=object)
    root = os.path.join(REF_ROOT, "openvis")
    _mod("refopenvis").__path__ = [root]
    _mod("refopenvis.modeling").__path__ = [os.path.join(root, "modeling")]
    _mod("refopenvis.utils").__path__ = [os.path.join(root, "utils")]
    _mod("refopenvis.modeling.transformer_decoder").__path__ = [_DEC_DIR]
    sys.modules["refopenvis.modeling.transformer_decoder.video_mask2former_transformer_decoder"] = dec.video
    _mod("refopenvis.modeling.video_maskformer", VideoMaskFormer=torch.nn.Module)
    index = _load("refopenvis.utils", os.path.join(root, "utils"), "index")
    minvis = _load("refopenvis.modeling", os.path.join(root, "modeling"), "minvis")
    resampler = _load("refopenvis.modeling", os.path.join(root, "modeling"), "resampler")
    ns = types.SimpleNamespace(match_via_embeds=minvis.match_via_embeds,
                               batch_video_match_via_embeds=minvis.batch_video_match_via_embeds,
                               batch_index=index.batch_index,
                               TemporalInstanceResampler=resampler.TemporalInstanceResampler)
    _loaded["temporal"] = ns
    return ns


def extract_function(path, qualname, env=None):
    """Compiles ONE function (or method: "Class.func") out of a reference source file without importing the module
    around it -- for code that sits in files whose imports (Detectron2 meta-architecture machinery, OpenAI clip) cannot
    be satisfied here.  The function body is the reference's own text, parsed with `ast` and executed in a namespace that
    holds only what it needs (`env`, default torch / F / List).  Nothing is copied into the repository."""
    import ast
    src = open(os.path.join(REF_ROOT, path)).read()
    tree = ast.parse(src)
    parts = qualname.split(".")
    body = tree.body
    node = None
    for i, name in enumerate(parts):
        node = next(n for n in body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name)
        body = getattr(node, "body", [])
    assert isinstance(node, ast.FunctionDef), qualname
    node.decorator_list = []
    mod = ast.Module(body=[node], type_ignores=[])
    ast.fix_missing_locations(mod)
    import typing
    import torch.nn.functional as F_
    ns = {"torch": torch, "F": F_, "List": typing.List, "nn": torch.nn}
    ns.update(env or {})
    exec(compile(mod, os.path.join(REF_ROOT, path), "exec"), ns)
    return ns[parts[-1]]


def ov_tails():
    """The open-vocabulary tails that live in files which cannot be imported here, as the reference's own functions:

      OpenVIS.open_vocabulary_inference     openvis/openvis.py:110-147      (row A15: per-query mean over valid frames, softmax)
      ClipAdapter.normalize / cal_sim_logits openvis/modeling/clip_adapter/adapter.py:118-119, 146-147
      ZeroShotClassifier.forward            openvis/ov2seg.py:515-529        (row A17: [text; 0] rows, scale 50)
    """
    if "ov" in _loaded:
        return _loaded["ov"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    ad = "openvis/modeling/clip_adapter/adapter.py"
    ns = types.SimpleNamespace(
        open_vocabulary_inference=extract_function("openvis/openvis.py", "OpenVIS.open_vocabulary_inference"),
        normalize=extract_function(ad, "ClipAdapter.normalize"),
        cal_sim_logits=extract_function(ad, "ClipAdapter.cal_sim_logits"),
        zero_shot_forward=extract_function("openvis/ov2seg.py", "ZeroShotClassifier.forward"))
    _loaded["ov"] = ns
    return ns


def zero_shot_decoder():
    """ZeroShotMultiScaleMaskedTransformerDecoder of the reference (zero_shot_mask2former_transformer_decoder.py:15-277)."""
    decoders()
    return _load("refdec", _DEC_DIR, "zero_shot_mask2former_transformer_decoder").ZeroShotMultiScaleMaskedTransformerDecoder


def embedding_decoders():
    """Embedding* / Proposal* decoder classes of the reference (video_..._decoder.py:487-537, frame_...:157-207)."""
    d = decoders()
    return types.SimpleNamespace(
        embedding_video=d.video.EmbeddingVideoMultiScaleMaskedTransformerDecoder,
        proposal_video=d.video.ProposalVideoMultiScaleMaskedTransformerDecoder,
        embedding_frame=d.frame.EmbeddingFrameMultiScaleMaskedTransformerDecoder,
        proposal_frame=d.frame.ProposalFrameMultiScaleMaskedTransformerDecoder,
        registry=d.video.TRANSFORMER_DECODER_REGISTRY,
        build_transformer_decoder=d.video.build_transformer_decoder)
