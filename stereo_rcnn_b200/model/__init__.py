"""Mirror of the reference's ``lib/model`` package tree for the hot path.

Put the directory that contains this package first on ``sys.path`` *as* ``model``
(``sys.path.insert(0, '<repo>/stereo_rcnn_b200')``) and the reference's scripts import
``model.nms.nms_wrapper.nms``, ``model.roi_align.modules.roi_align.RoIAlignAvg``,
``model.rpn.proposal_layer._ProposalLayer``, ``model.dense_align.dense_align.align_parallel``
and ``model.stereo_rcnn.resnet.resnet`` with unchanged names and call signatures -- see
INTEGRATION.md.  Inside this repo the same modules are ``stereo_rcnn_b200.model.*``.
"""
