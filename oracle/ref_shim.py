"""Run the reference's own Python (lib/model/...) under py3.12 / torch 2.x on CPU.

TEST INFRASTRUCTURE ONLY, and only usable where ``/root/reference`` exists (the
build container).  Nothing is copied: the reference sources are read from where
they lie, a handful of *textual* py2->py3 / torch-0.3->2.x patches are applied
in memory, and the modules are exec'd into ``sys.modules`` under their original
names (``model.rpn.proposal_layer`` ...).  The two native extensions the
reference needs (cffi NMS / RoIAlign, which cannot be built on this stack:
``torch.utils.ffi`` and THC are gone) are replaced by the C oracle, which is
itself pinned to the reference ``.cu`` files on the GPU box (oracle/_ref).

Shims:
  * ``easydict`` stub;  ``.cuda()`` -> identity, ``torch.cuda.FloatTensor`` -> CPU
  * ``F.upsample`` / ``F.grid_sample`` pinned to ``align_corners=True`` (torch 0.3.0
    semantics, Q4/Q20)
  * legacy ``torch.cat`` rank padding in box_3d.py:97 (Q21)
  * implicit relative imports and the py2 ``print`` block of generate_anchors.py
  * train-only target layers stubbed (out of this round's scope)
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

REF = os.environ.get("STEREO_REFERENCE", "/root/reference")
LIB = os.path.join(REF, "lib")


def available():
    return os.path.isdir(os.path.join(LIB, "model"))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)
        super().__setattr__(k, v)

    __setitem__ = __setattr__


def _exec_module(name, path, patches=(), is_pkg=False):
    with open(path) as f:
        src = f.read()
    for old, new in patches:
        assert old in src, "shim patch target missing in %s: %r" % (path, old)
        src = src.replace(old, new)
    mod = types.ModuleType(name)
    mod.__file__ = path
    if is_pkg:
        mod.__path__ = [os.path.dirname(path)]
    mod.__package__ = name if is_pkg else name.rsplit(".", 1)[0]
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


_loaded = None


def load():
    """returns a namespace with the reference's hot-path Python, runnable on CPU"""
    global _loaded
    if _loaded is not None:
        return _loaded
    assert available(), "reference checkout not found at %s" % REF

    # --- environment shims ---------------------------------------------------
    ed = types.ModuleType("easydict")
    ed.EasyDict = _EasyDict
    sys.modules.setdefault("easydict", ed)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    _interp = F.interpolate

    def _upsample(x, size=None, scale_factor=None, mode="nearest", align_corners=None):
        return _interp(x, size=size, scale_factor=scale_factor, mode=mode,
                       align_corners=True if mode == "bilinear" else None)

    F.upsample = _upsample
    _gs = F.grid_sample

    def _grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
        return _gs(inp, grid, mode=mode, padding_mode=padding_mode, align_corners=True)

    F.grid_sample = _grid_sample

    # --- package skeleton ------------------------------------------------------
    if LIB not in sys.path:
        sys.path.insert(0, LIB)
    for pkg in ("model", "model.utils", "model.rpn", "model.nms", "model.roi_align",
                "model.roi_align.modules", "model.dense_align", "model.stereo_rcnn"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(LIB, *pkg.split("."))]
            sys.modules[pkg] = m

    cfgmod = _exec_module("model.utils.config", os.path.join(LIB, "model/utils/config.py"))
    cfg = cfgmod.cfg

    # native-extension stand-ins (C oracle)
    nmsw = types.ModuleType("model.nms.nms_wrapper")

    def nms(dets, thresh, force_cpu=False):
        if dets.shape[0] == 0:
            return []
        keep = ops.nms(dets.detach().cpu().numpy(), float(thresh))
        return torch.from_numpy(keep.astype(np.int32)).view(-1, 1)

    nmsw.nms = nms
    sys.modules["model.nms.nms_wrapper"] = nmsw

    ram = types.ModuleType("model.roi_align.modules.roi_align")

    class RoIAlignAvg(nn.Module):
        def __init__(self, aligned_height, aligned_width, spatial_scale):
            super().__init__()
            self.aligned_width, self.aligned_height = int(aligned_width), int(aligned_height)

        def forward(self, features, rois, scale):
            r = rois.detach().numpy().reshape(-1, 5)
            out = ops.roi_align_avg(features.detach().numpy(), r, self.aligned_height,
                                    self.aligned_width, np.float32(float(scale)))
            return torch.from_numpy(out)

    ram.RoIAlignAvg = RoIAlignAvg
    sys.modules["model.roi_align.modules.roi_align"] = ram

    for name, cls in (("model.rpn.anchor_target_layer", "_AnchorTargetLayer"),
                      ("model.rpn.proposal_target_layer", "_ProposalTargetLayer")):
        m = types.ModuleType(name)
        setattr(m, cls, type(cls, (nn.Module,), {"__init__": lambda self, *a, **k: nn.Module.__init__(self)}))
        sys.modules[name] = m

    # --- reference modules, patched in memory ------------------------------------
    bt = _exec_module("model.rpn.bbox_transform", os.path.join(LIB, "model/rpn/bbox_transform.py"))
    with open(os.path.join(LIB, "model/rpn/generate_anchors.py")) as f:
        ga_src = f.read()
    head, tail = ga_src.split("if __name__ == '__main__':", 1)
    tail = tail.split("############################################################", 1)[1]
    ga = types.ModuleType("model.rpn.generate_anchors")
    sys.modules["model.rpn.generate_anchors"] = ga
    exec(compile(head + "\n####" + tail, "generate_anchors.py", "exec"), ga.__dict__)

    pl = _exec_module("model.rpn.proposal_layer", os.path.join(LIB, "model/rpn/proposal_layer.py"), [
        ("from generate_anchors import", "from model.rpn.generate_anchors import"),
        ("from bbox_transform import", "from model.rpn.bbox_transform import"),
    ])
    nu = _exec_module("model.utils.net_utils", os.path.join(LIB, "model/utils/net_utils.py"))
    rpn = _exec_module("model.rpn.stereo_rpn", os.path.join(LIB, "model/rpn/stereo_rpn.py"), [
        ("from proposal_layer import", "from model.rpn.proposal_layer import"),
        ("from anchor_target_layer import", "from model.rpn.anchor_target_layer import"),
    ])
    srcnn = _exec_module("model.stereo_rcnn.stereo_rcnn",
                         os.path.join(LIB, "model/stereo_rcnn/stereo_rcnn.py"), [
        ("idx_l = (roi_level == l).nonzero().squeeze()", "idx_l = (roi_level == l).nonzero().view(-1)"),
    ])
    resnet = _exec_module("model.stereo_rcnn.resnet", os.path.join(LIB, "model/stereo_rcnn/resnet.py"))
    ku = types.ModuleType("model.utils.kitti_utils")       # box_3d.py imports it but never uses it
    sys.modules["model.utils.kitti_utils"] = ku
    sys.modules["model.utils"].kitti_utils = ku
    b3 = _exec_module("model.dense_align.box_3d", os.path.join(LIB, "model/dense_align/box_3d.py"), [
        ("torch.cat((pt2, torch.ones_like(pt2[:,:,0])),2)",
         "torch.cat((pt2, torch.ones_like(pt2[:,:,0:1])),2)"),
    ])
    da = _exec_module("model.dense_align.dense_align",
                      os.path.join(LIB, "model/dense_align/dense_align.py"))

    _loaded = types.SimpleNamespace(cfg=cfg, bbox_transform=bt, generate_anchors=ga,
                                    proposal_layer=pl, stereo_rpn=rpn, stereo_rcnn=srcnn,
                                    resnet=resnet, box_3d=b3, dense_align=da, net_utils=nu)
    return _loaded


_loaded_train = None


def load_train():
    """the reference's REAL train-time target layers (anchor_target_layer.py, proposal_target_layer.py), exec'd on CPU
    with py2 -> py3 / torch 0.3 -> 2.x text patches only (`long(` -> `int(`, the removed `Tensor.index`, implicit
    relative imports); `load()` keeps stubs under the module names because the eval-mode model never calls them."""
    global _loaded_train
    if _loaded_train is not None:
        return _loaded_train
    ref = load()
    torch.cuda.LongTensor = torch.LongTensor
    atl = _exec_module("model.rpn._anchor_target_layer_real", os.path.join(LIB, "model/rpn/anchor_target_layer.py"), [
        ("from generate_anchors import", "from model.rpn.generate_anchors import"),
        ("from bbox_transform import", "from model.rpn.bbox_transform import"),
        ("long(im_info[0][1])", "int(im_info[0][1])"),
        ("long(im_info[0][0])", "int(im_info[0][0])"),
    ])
    ptl = _exec_module("model.rpn._proposal_target_layer_real", os.path.join(LIB, "model/rpn/proposal_target_layer.py"), [
        ("from ..utils.config import cfg", "from model.utils.config import cfg"),
        ("from bbox_transform import", "from model.rpn.bbox_transform import"),
        (".contiguous().view(-1).index(offset.view(-1))", ".contiguous().view(-1)[offset.view(-1)]"),
    ])
    _loaded_train = types.SimpleNamespace(cfg=ref.cfg, anchor_target_layer=atl, proposal_target_layer=ptl,
                                          net_utils=ref.net_utils, generate_anchors=ref.generate_anchors)
    return _loaded_train


class Calib(object):
    """minimal stand-in for kitti_utils.FrameCalibrationData (only p2/p3 are read)"""

    def __init__(self, p2, p3):
        self.p2, self.p3 = np.asarray(p2, np.float64), np.asarray(p3, np.float64)


def demo_calib():
    """P2 / P3 of the reference's demo/calib.txt"""
    rows = {}
    with open(os.path.join(REF, "demo/calib.txt")) as f:
        for line in f:
            if ":" in line:
                k, v = line.split(":", 1)
                rows[k.strip()] = np.array([float(t) for t in v.split()], np.float64)
    return Calib(rows["P2"].reshape(3, 4), rows["P3"].reshape(3, 4))


def build_reference_model(state_dict, classes=("__background__", "Car")):
    """reference ``resnet(classes,101).create_architecture()`` in eval mode with our weights"""
    ref = load()
    torch.manual_seed(0)
    m = ref.resnet.resnet(classes, 101, pretrained=False)
    m.create_architecture()
    missing, unexpected = m.load_state_dict(state_dict, strict=False)
    unexpected = [k for k in unexpected]
    missing = [k for k in missing if not k.endswith("num_batches_tracked")]
    assert not missing and not unexpected, (missing, unexpected)
    m.eval()
    return m
