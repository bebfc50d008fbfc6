// nms.cu -- bitmask NMS with a device-side greedy reduction (sm_100a).
//
// Replaces lib/model/nms/src/nms_cuda_kernel.cu:41-161 of the reference
// (nms_kernel + the host-side scan of nms_cuda_compute).  Same arithmetic
// (devIoU with "+1" areas, strict '>' against the threshold, boxes pre-sorted),
// different machine mapping:
//   * only the upper triangle of 64x64 tiles is computed (the reference computes
//     the full square and never reads the lower half);
//   * the greedy scan runs on the device, one 64-box chunk at a time: the chunk's
//     diagonal tile is resolved sequentially from shared memory, then every kept
//     row of the chunk is OR-ed into the running suppression words by one thread
//     per 64-bit word (coalesced row reads, independent loads in flight);
//   * no cudaMalloc / cudaMemcpy / default-stream sync: workspace is caller-owned,
//     everything is ordered on the caller's stream;
//   * two box sets that share scores (left / right proposals) are reduced in
//     lock step and intersected on the fly (np.intersect1d of
//     proposal_layer.py:128), stopping as soon as `max_out` survivors exist.
#include "common.cuh"

namespace {

constexpr int kTile = 64;
constexpr int kMaxWords = 256;  // n <= 16384

__global__ void __launch_bounds__(kTile)
nms_mask_kernel(const float* __restrict__ boxes0, const float* __restrict__ boxes1, int stride,
                int n, float thresh, unsigned long long* __restrict__ mask0,
                unsigned long long* __restrict__ mask1) {
    const int row = blockIdx.y, col = blockIdx.x;
    if (col < row) return;  // upper triangle only
    const float* boxes = blockIdx.z ? boxes1 : boxes0;
    unsigned long long* mask = blockIdx.z ? mask1 : mask0;
    const int cb = (n + kTile - 1) / kTile;
    const int row_size = min(n - row * kTile, kTile);
    const int col_size = min(n - col * kTile, kTile);
    __shared__ float4 cbox[kTile];
    const int t = threadIdx.x;
    if (t < col_size) {
        const float* p = boxes + (size_t)(col * kTile + t) * stride;
        cbox[t] = make_float4(p[0], p[1], p[2], p[3]);
    }
    __syncthreads();
    if (t < row_size) {
        const int i = row * kTile + t;
        const float* p = boxes + (size_t)i * stride;
        const float4 cur = make_float4(p[0], p[1], p[2], p[3]);
        unsigned long long bits = 0;
        const int start = (row == col) ? t + 1 : 0;
        for (int j = start; j < col_size; ++j)
            if (sb_iou_gt(cur, cbox[j], thresh)) bits |= 1ULL << j;
        mask[(size_t)i * cb + col] = bits;
    }
}

// One CTA, 256 threads per side.  keep_out receives ascending indices of boxes kept by
// *every* side; at most max_out of them.
template <int NS>
__global__ void __launch_bounds__(256 * NS)
nms_reduce_kernel(const unsigned long long* __restrict__ mask0,
                  const unsigned long long* __restrict__ mask1, int n, int max_out,
                  int* __restrict__ keep_out, int* __restrict__ num_out) {
    __shared__ unsigned long long remv[NS][kMaxWords];
    __shared__ unsigned long long keepbits[NS];
    __shared__ int count;
    const int side = threadIdx.x >> 8, t = threadIdx.x & 255;
    const unsigned long long* mask = side ? mask1 : mask0;
    const int cb = (n + kTile - 1) / kTile;
    for (int w = t; w < kMaxWords; w += 256) remv[side][w] = 0;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    for (int c = 0; c < cb; ++c) {
        const int csize = min(n - c * kTile, kTile);
        // resolve the chunk's 64x64 diagonal tile inside warp 0 of each side: lane l holds rows l and l+32 in
        // registers and the sequential scan broadcasts row b with a shuffle (no shared-memory latency on the
        // 64-step dependency chain)
        if (t < 32) {
            const int r0 = c * kTile + t, r1 = r0 + 32;
            const unsigned long long d0 = (t < csize) ? mask[(size_t)r0 * cb + c] : 0ULL;
            const unsigned long long d1 = (t + 32 < csize) ? mask[(size_t)r1 * cb + c] : 0ULL;
            unsigned long long cur = remv[side][c], kb = 0;
            const unsigned long long valid = csize >= 64 ? ~0ULL : ((1ULL << csize) - 1ULL);
            cur |= ~valid;
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                const unsigned long long row = __shfl_sync(0xffffffffu, d0, b);
                if (!((cur >> b) & 1ULL)) { kb |= 1ULL << b; cur |= row; }
            }
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                const unsigned long long row = __shfl_sync(0xffffffffu, d1, b);
                if (!((cur >> (b + 32)) & 1ULL)) { kb |= 1ULL << (b + 32); cur |= row; }
            }
            if (t == 0) keepbits[side] = kb;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long both = keepbits[0];
            if (NS == 2) both &= keepbits[NS - 1];
            int cnt = count;
            while (both && cnt < max_out) {
                int b = __ffsll((long long)both) - 1;
                both &= both - 1;
                keep_out[cnt++] = c * kTile + b;
            }
            count = cnt;
        }
        {
            // OR the rows of this side's kept boxes into the suppression words beyond chunk c
            unsigned long long kb = keepbits[side];
            for (int w = t; w < cb; w += 256) {
                if (w <= c) continue;
                unsigned long long acc = 0;
                unsigned long long k2 = kb;
                while (k2) {
                    int b = __ffsll((long long)k2) - 1;
                    k2 &= k2 - 1;
                    acc |= mask[(size_t)(c * kTile + b) * cb + w];
                }
                remv[side][w] |= acc;
            }
        }
        __syncthreads();
        if (count >= max_out) break;
    }
    if (threadIdx.x == 0) *num_out = count;
}

}  // namespace

// shared with proposal.cu
int sb_nms_launch(const float* boxes0, const float* boxes1, int stride, int n, float thresh,
                  unsigned long long* mask0, unsigned long long* mask1, int max_out, int* keep,
                  int* num_out, cudaStream_t st) {
    if (n > kTile * kMaxWords) return SB_EINVAL;
    const int ns = boxes1 ? 2 : 1;
    if (n == 0) {
        cudaMemsetAsync(num_out, 0, sizeof(int), st);
        return SB_OK;
    }
    const int cb = (n + kTile - 1) / kTile;
    dim3 grid(cb, cb, ns);
    nms_mask_kernel<<<grid, kTile, 0, st>>>(boxes0, boxes1, stride, n, thresh, mask0, mask1);
    SB_LAUNCHED();
    SB_CHECK_LAUNCH();
    if (ns == 1)
        nms_reduce_kernel<1><<<1, 256, 0, st>>>(mask0, mask0, n, max_out, keep, num_out);
    else
        nms_reduce_kernel<2><<<1, 512, 0, st>>>(mask0, mask1, n, max_out, keep, num_out);
    SB_LAUNCHED();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" size_t sb_nms_workspace_bytes(int n) {
    size_t cb = (size_t)(n + kTile - 1) / kTile;
    return (size_t)n * cb * sizeof(unsigned long long) + 256;
}

extern "C" int sb_nms(const float* dets, int n, float thresh, int* keep, int* num_out,
                      void* workspace, size_t workspace_bytes, sb_stream_t stream) {
    if (n < 0 || !keep || !num_out) return SB_EINVAL;
    if (n > 0 && (!dets || !workspace || workspace_bytes < sb_nms_workspace_bytes(n))) return SB_EINVAL;
    return sb_nms_launch(dets, nullptr, 5, n, thresh, (unsigned long long*)workspace, nullptr, n, keep,
                         num_out, sb_cs(stream));
}

extern "C" int sb_nms_mask(const float* dets, int n, float thresh, uint64_t* mask, sb_stream_t stream) {
    if (n <= 0) return n == 0 ? SB_OK : SB_EINVAL;
    if (n > kTile * kMaxWords) return SB_EINVAL;
    const int cb = (n + kTile - 1) / kTile;
    cudaMemsetAsync(mask, 0, (size_t)n * cb * sizeof(uint64_t), sb_cs(stream));
    dim3 grid(cb, cb, 1);
    nms_mask_kernel<<<grid, kTile, 0, sb_cs(stream)>>>(dets, nullptr, 5, n, thresh,
                                                      (unsigned long long*)mask, nullptr);
    SB_LAUNCHED();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
