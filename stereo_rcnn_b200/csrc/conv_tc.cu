// conv_tc.cu -- placeholder until the tcgen05 kernel lands (next commit).
#include "common.cuh"
extern "C" int sb_conv2d_tc_supported(const sb_conv_desc* d) { (void)d; return 0; }
extern "C" int sb_conv2d_tc(const sb_conv_desc* d, sb_stream_t stream) { (void)d; (void)stream; return SB_EINVAL; }
