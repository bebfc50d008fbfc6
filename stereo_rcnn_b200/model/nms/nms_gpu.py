"""lib/model/nms/nms_gpu.py:7-12 -- caller allocates outputs, native code fills them."""
from ._ext import nms


def nms_gpu(dets, thresh):
    keep = dets.new_zeros((dets.size(0), 1)).int()
    num_out = dets.new_zeros(1).int()
    nms.nms_cuda(keep, dets, num_out, thresh)
    return keep[:int(num_out[0])]
