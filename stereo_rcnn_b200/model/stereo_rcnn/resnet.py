"""``resnet(classes, 101, pretrained).create_architecture()`` with the reference's public interface
(lib/model/stereo_rcnn/resnet.py:220-348 + stereo_rcnn.py:141-324): same ``state_dict`` keys
(``RCNN_layer1.0.0.conv1.weight``, ``RCNN_rpn.RPN_Conv.weight`` ...), same forward argument list and
15-tuple result, so test_net.py / demo.py keep working.  The modules below only *hold* parameters;
``forward`` hands them to the sm_100a engine (stereo_rcnn_b200.engine) -- no torch.nn op runs.

Training mode (trainval_net.py) raises NotImplementedError here: the target layers, the losses and their gradients
w.r.t. the network outputs exist on the device (stereo_rcnn_b200.train, model.rpn.anchor_target_layer /
proposal_target_layer), the backward pass of the trunk and heads (dgrad / wgrad) does not.
"""
import math

import torch
import torch.nn as nn

from stereo_rcnn_b200 import engine as _engine
from ..utils.config import cfg

_LAYERS = [3, 4, 23, 3]
_PLANES = [64, 128, 256, 512]


def _bottleneck(inplanes, planes, stride, with_ds):
    m = nn.Module()
    m.conv1 = nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False)     # stride on the 1x1 (Q1)
    m.bn1 = nn.BatchNorm2d(planes)
    m.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=False)
    m.bn2 = nn.BatchNorm2d(planes)
    m.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
    m.bn3 = nn.BatchNorm2d(planes * 4)
    if with_ds:
        m.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                     nn.BatchNorm2d(planes * 4))
    return m


class _RPNParams(nn.Module):
    def __init__(self, din):
        super().__init__()
        self.RPN_Conv = nn.Conv2d(din, 512, 3, 1, 1, bias=True)
        self.RPN_cls_score = nn.Conv2d(1024, 6, 1, 1, 0)
        self.RPN_bbox_pred_left_right = nn.Conv2d(1024, 18, 1, 1, 0)


class resnet(nn.Module):
    def __init__(self, classes, num_layers=101, pretrained=False):
        super().__init__()
        assert num_layers == 101, "the Stereo R-CNN hot path is ResNet-101"
        self.classes = classes
        self.n_classes = len(classes)
        self.model_path = 'data/pretrained_model/resnet101_caffe.pth'
        self.dout_base_model = 256
        self.pretrained = pretrained
        self._engine = None

    # ---------------------------------------------------------------- construction
    def _init_modules(self):
        self.RCNN_layer0 = nn.Sequential(nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64),
                                         nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 0, ceil_mode=True))
        inpl = 64
        for li, (nb, pl) in enumerate(zip(_LAYERS, _PLANES)):
            blocks = []
            for b in range(nb):
                blocks.append(_bottleneck(inpl, pl, (2 if li > 0 else 1) if b == 0 else 1, b == 0))
                inpl = pl * 4
            setattr(self, "RCNN_layer%d" % (li + 1), nn.Sequential(nn.Sequential(*blocks)))
        self.RCNN_toplayer = nn.Conv2d(2048, 256, 1)
        self.RCNN_smooth1 = nn.Conv2d(256, 256, 3, padding=1)
        self.RCNN_smooth2 = nn.Conv2d(256, 256, 3, padding=1)
        self.RCNN_smooth3 = nn.Conv2d(256, 256, 3, padding=1)
        self.RCNN_latlayer1 = nn.Conv2d(1024, 256, 1)
        self.RCNN_latlayer2 = nn.Conv2d(512, 256, 1)
        self.RCNN_latlayer3 = nn.Conv2d(256, 256, 1)
        self.RCNN_rpn = _RPNParams(self.dout_base_model)
        self.RCNN_top = nn.Sequential(nn.Conv2d(512, 2048, cfg.POOLING_SIZE, stride=cfg.POOLING_SIZE), nn.ReLU(True),
                                      nn.Dropout(p=0.2), nn.Conv2d(2048, 2048, 1), nn.ReLU(True), nn.Dropout(p=0.2))
        k = []
        for _ in range(6):
            k += [nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(True)]
        k += [nn.ConvTranspose2d(256, 256, 2, stride=2), nn.ReLU(True)]
        self.RCNN_kpts = nn.Sequential(*k)
        self.RCNN_cls_score = nn.Linear(2048, self.n_classes)
        self.RCNN_bbox_pred = nn.Linear(2048, 6 * self.n_classes)
        self.RCNN_dim_orien_pred = nn.Linear(2048, 5 * self.n_classes)
        self.kpts_class = nn.Conv2d(256, 6, 1)
        if self.pretrained:
            sd = torch.load(self.model_path)
            own = self.state_dict()
            remap = {"conv1.": "RCNN_layer0.0.", "bn1.": "RCNN_layer0.1."}
            for key, v in sd.items():
                for old, new in remap.items():
                    if key.startswith(old):
                        key = new + key[len(old):]
                if key.startswith("layer"):
                    key = "RCNN_" + key[:6] + ".0" + key[6:]
                if key in own:
                    own[key].copy_(v)
        for p in self.parameters():
            p.requires_grad = False

    def _init_weights(self):
        """stereo_rcnn.py:47-85 + resnet.py:123-129 initial distributions"""
        for m in self.modules():
            if isinstance(m, nn.Conv2d) and m.bias is None:
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))

        def normal_init(m, std):
            m.weight.data.normal_(0, std)
            m.bias.data.zero_()
        for m in (self.RCNN_toplayer, self.RCNN_smooth1, self.RCNN_smooth2, self.RCNN_smooth3, self.RCNN_latlayer1,
                  self.RCNN_latlayer2, self.RCNN_latlayer3, self.RCNN_rpn.RPN_Conv, self.RCNN_rpn.RPN_cls_score,
                  self.RCNN_rpn.RPN_bbox_pred_left_right, self.RCNN_cls_score):
            normal_init(m, 0.01)
        normal_init(self.RCNN_bbox_pred, 0.001)
        normal_init(self.RCNN_dim_orien_pred, 0.001)
        normal_init(self.kpts_class, 0.1)
        for seq in (self.RCNN_top, self.RCNN_kpts):
            for m in seq:
                if hasattr(m, "weight"):
                    normal_init(m, 0.02)

    def create_architecture(self):
        self._init_modules()
        self._init_weights()

    def load_state_dict(self, state_dict, strict=True):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict)

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("stereo_rcnn_b200: no backward pass of the network (dgrad / wgrad); target layers and "
                                      "losses are in stereo_rcnn_b200.train")
        return super().train(False)

    # ---------------------------------------------------------------- forward
    def forward(self, im_left_data, im_right_data, im_info, gt_boxes_left=None, gt_boxes_right=None,
                gt_boxes_merge=None, gt_dim_orien=None, gt_kpts=None, num_boxes=None):
        if self._engine is None:
            self._engine = _engine.StereoRCNNEngine(self.state_dict(), device=im_left_data.device,
                                                    n_classes=self.n_classes)
        o = self._engine.forward(im_left_data, im_right_data, im_info, "TEST")
        zero = 0
        return (o["rois_left"], o["rois_right"], o["cls_prob"], o["bbox_pred"], o["dim_orien_pred"],
                o["kpts_prob"], o["left_border_prob"], o["right_border_prob"], zero, zero, zero, zero, zero, zero,
                None)
