// train_targets.cu -- the train-time target layers of the reference on the device, no host
// round trip (the reference walks Python loops with per-element GPU indexing and `.cpu()` copies):
//
//   sb_anchor_targets    lib/model/rpn/anchor_target_layer.py:42-164   RPN labels / box targets / loss weights
//   sb_proposal_targets  lib/model/rpn/proposal_target_layer.py:36-333 sampled RoIs + box / dimension / keypoint targets
//
// Arithmetic: every fp32 operation of bbox_overlaps_batch (bbox_transform.py:220-309) and bbox_transform_batch
// (:38-77) is a single IEEE operation in the reference's order (_rn intrinsics, no FMA contraction), so every compare
// against a threshold -- labels, foreground / background sets, sampled indices -- is bit-identical to the CPU oracle
// (oracle/train_targets.py, itself pinned to the reference's layers); only `log` may differ in the last ulp.
//
// Sampling: the reference consumes numpy's global stream with data-dependent lengths (np.random.permutation(n_fg),
// np.random.rand(k)), which cannot be reproduced without copying counts to the host.  The kernels take explicit
// random words instead (oracle: KeySampler): a key per anchor / RoI -- "a random permutation of the candidates" is
// their order by (key, index) -- and one word per background draw (index = word * n >> 32).  Everything else is the
// reference's algorithm including its quirks (see oracle/train_targets.py:anchor_target_layer).
#include "common.cuh"

namespace {

constexpr int kMaxGt = 64;            // cfg.MAX_NUM_GT_BOXES = 30

__device__ __forceinline__ float box_w(float x1, float x2) { return __fadd_rn(__fsub_rn(x2, x1), 1.0f); }

// one entry of bbox_overlaps_batch (before the zero-area masks)
__device__ __forceinline__ float overlap(const float4 a, const float aarea, const float4 g, const float garea) {
    float iw = __fadd_rn(__fsub_rn(fminf(a.z, g.z), fmaxf(a.x, g.x)), 1.0f);
    if (iw < 0.f) iw = 0.f;
    float ih = __fadd_rn(__fsub_rn(fminf(a.w, g.w), fmaxf(a.y, g.y)), 1.0f);
    if (ih < 0.f) ih = 0.f;
    const float inter = __fmul_rn(iw, ih);
    const float ua = __fsub_rn(__fadd_rn(aarea, garea), inter);
    return __fdiv_rn(inter, ua);
}

struct GtSet {                        // the K ground-truth boxes of one image, staged in shared memory
    float4 box[kMaxGt];
    float area[kMaxGt];
    unsigned char zero[kMaxGt];
};

__device__ __forceinline__ void load_gt(GtSet& s, const float* __restrict__ gt /*[K][5]*/, int K) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float4 g = make_float4(gt[k * 5 + 0], gt[k * 5 + 1], gt[k * 5 + 2], gt[k * 5 + 3]);
        const float gx = box_w(g.x, g.z), gy = box_w(g.y, g.w);
        s.box[k] = g;
        s.area[k] = __fmul_rn(gx, gy);
        s.zero[k] = (gx == 1.f && gy == 1.f) ? 1 : 0;
    }
}

// overlaps of one box against the staged set: max and FIRST argmax (torch.max on CPU), with the masks of
// bbox_transform.py:261-262 (zero-area gt -> 0, zero-area anchor -> -1)
__device__ __forceinline__ void max_overlap(const GtSet& s, int K, const float4 a, float& best, int& arg) {
    const float ax = box_w(a.x, a.z), ay = box_w(a.y, a.w);
    const float aarea = __fmul_rn(ax, ay);
    const bool azero = ax == 1.f && ay == 1.f;
    best = -INFINITY;
    arg = 0;
    for (int k = 0; k < K; ++k) {
        float ov = overlap(a, aarea, s.box[k], s.area[k]);
        if (s.zero[k]) ov = 0.f;
        if (azero) ov = -1.f;
        if (ov > best) { best = ov; arg = k; }
    }
}

// order-preserving map of a float onto unsigned (for atomicMax)
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bbox_transform_batch for one (ex, gt) pair (bbox_transform.py:57-70)
__device__ __forceinline__ float4 box_delta(const float4 ex, const float4 gt) {
    const float ew = box_w(ex.x, ex.z), eh = box_w(ex.y, ex.w);
    const float ecx = __fadd_rn(ex.x, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(ex.y, __fmul_rn(0.5f, eh));
    const float gw = box_w(gt.x, gt.z), gh = box_w(gt.y, gt.w);
    const float gcx = __fadd_rn(gt.x, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(gt.y, __fmul_rn(0.5f, gh));
    return make_float4(__fdiv_rn(__fsub_rn(gcx, ecx), ew), __fdiv_rn(__fsub_rn(gcy, ecy), eh),
                       logf(__fdiv_rn(gw, ew)), logf(__fdiv_rn(gh, eh)));
}

__device__ __forceinline__ bool anchor_inside(const float4 a, float im_w, float im_h) {
    return a.x >= 0.f && a.y >= 0.f && a.z < im_w && a.w < im_h;            // anchor_target_layer.py:70-73, border 0
}

// ------------------------------------------------------------------------------------------ anchor targets
// pass 1: per inside anchor max / argmax overlap; per gt box the maximum over the inside anchors
template <int KC>                                     // compile-time cap of K (32 covers cfg.MAX_NUM_GT_BOXES = 30)
__global__ void __launch_bounds__(256)
at_overlap_kernel(const float4* __restrict__ anchors, int A, const float* __restrict__ gt_merge, int K, float im_w,
                  float im_h, float* __restrict__ max_ov, int* __restrict__ arg_ov, unsigned* __restrict__ gt_max) {
    __shared__ GtSet s;
    __shared__ unsigned s_max[kMaxGt];
    const int b = blockIdx.y;
    load_gt(s, gt_merge + (size_t)b * K * 5, K);
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_max[k] = 0u;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned ovk[KC];                                            // ordered overlap per gt box; 0 = not a candidate
#pragma unroll
    for (int k = 0; k < KC; ++k) ovk[k] = 0u;
    if (i < A) {
        const float4 a = anchors[i];
        float best = -1.f;
        int arg = -1;
        if (anchor_inside(a, im_w, im_h)) {
            const float ax = box_w(a.x, a.z), ay = box_w(a.y, a.w);
            const float aarea = __fmul_rn(ax, ay);
            const bool azero = ax == 1.f && ay == 1.f;
            best = -INFINITY;
            arg = 0;
#pragma unroll
            for (int k = 0; k < KC; ++k) {                      // unrolled: ovk[] stays in registers
                if (k < K) {
                    float ov = overlap(a, aarea, s.box[k], s.area[k]);
                    if (s.zero[k]) ov = 0.f;
                    if (azero) ov = -1.f;
                    if (ov > best) { best = ov; arg = k; }
                    ovk[k] = f2ord(ov);
                }
            }
        }
        max_ov[(size_t)b * A + i] = best;
        arg_ov[(size_t)b * A + i] = arg;                         // -1: outside the image
    }
    // per-gt maximum: warp reduction, one shared atomic per warp and box, one global atomic per CTA and box
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        if (k < K) {
            const unsigned m = __reduce_max_sync(0xffffffffu, ovk[k]);
            if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_max[k], m);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (s_max[k]) atomicMax(&gt_max[b * kMaxGt + k], s_max[k]);
}

// pass 2: labels before subsampling (anchor_target_layer.py:85-101) + foreground / background counts
__global__ void __launch_bounds__(256)
at_label_kernel(const float4* __restrict__ anchors, int A, const float* __restrict__ gt_merge, int K,
                const float* __restrict__ max_ov, const int* __restrict__ arg_ov, const unsigned* __restrict__ gt_max,
                float neg_thr, float pos_thr, float* __restrict__ labels, int* __restrict__ counts /*[B][2]*/) {
    __shared__ GtSet s;
    __shared__ float s_gmax[kMaxGt];
    __shared__ int s_cnt[2];
    const int b = blockIdx.y;
    load_gt(s, gt_merge + (size_t)b * K * 5, K);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const unsigned u = gt_max[b * kMaxGt + k];
        float g = u ? ord2f(u) : -INFINITY;                      // no inside anchor at all: nothing can match
        if (g == 0.f) g = 1e-5f;                                 // :89
        s_gmax[k] = g;
    }
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A) {
        float lab = -1.f;
        if (arg_ov[(size_t)b * A + i] >= 0) {
            const float4 a = anchors[i];
            const float mx = max_ov[(size_t)b * A + i];
            if (mx < neg_thr) lab = 0.f;
            const float ax = box_w(a.x, a.z), ay = box_w(a.y, a.w);
            const float aarea = __fmul_rn(ax, ay);
            const bool azero = ax == 1.f && ay == 1.f;
            bool hit = false;
            for (int k = 0; k < K; ++k) {
                float ov = overlap(a, aarea, s.box[k], s.area[k]);
                if (s.zero[k]) ov = 0.f;
                if (azero) ov = -1.f;
                hit |= ov == s_gmax[k];
            }
            if (hit) lab = 1.f;
            if (mx >= pos_thr) lab = 1.f;
            if (lab == 1.f) atomicAdd(&s_cnt[0], 1);
            if (lab == 0.f) atomicAdd(&s_cnt[1], 1);
        }
        labels[(size_t)b * A + i] = lab;
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&counts[b * 2 + threadIdx.x], s_cnt[threadIdx.x]);
}

// pass 3: subsampling (:103-123).  One CTA per (image, side); side 0 = foreground, 1 = background.  "Disable the
// first n - keep entries of a random permutation of the candidates" = disable the n - keep candidates with the
// smallest (key, index): a 4-pass radix select on the key, ties at the threshold resolved in index order.
constexpr int kSelThreads = 1024;
constexpr int kSelUnroll = 8;

__global__ void __launch_bounds__(kSelThreads)
at_sample_kernel(float* __restrict__ labels, const unsigned* __restrict__ keys, int A, const int* __restrict__ counts,
                 int rpn_batch, int num_fg) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_remaining, s_bucket;
    __shared__ int s_scan[kSelThreads / 32];
    __shared__ int s_base;
    const int b = blockIdx.x, side = blockIdx.y;
    const int sum_fg = counts[b * 2 + 0], sum_bg = counts[b * 2 + 1];
    const int n = side == 0 ? sum_fg : sum_bg;
    const int keep = side == 0 ? num_fg : rpn_batch - sum_fg;                // :114 uses the count BEFORE fg sampling
    if (n <= keep) return;                                                  // CTA-uniform
    const long long want = (long long)n - keep;                             // how many to disable (> n if keep < 0)
    float* lab = labels + (size_t)b * A;
    const unsigned* key = keys + (size_t)b * A;
    const float mine = side == 0 ? 1.f : 0.f;
    if (want >= n) {                                                        // everything goes
        for (int i = threadIdx.x; i < A; i += blockDim.x)
            if (lab[i] == mine) lab[i] = -1.f;
        return;
    }
    unsigned d = (unsigned)want;                  // find T = the d-th smallest key (1-based) among the candidates
    unsigned prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int base = 0; base < A; base += kSelThreads * kSelUnroll) {      // kSelUnroll independent loads in flight
            float lv[kSelUnroll];
            unsigned kv[kSelUnroll];
#pragma unroll
            for (int u = 0; u < kSelUnroll; ++u) {
                const int i = base + u * kSelThreads + threadIdx.x;
                lv[u] = i < A ? lab[i] : -2.f;
                kv[u] = i < A ? key[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kSelUnroll; ++u)
                if (lv[u] == mine && (kv[u] & mask) == prefix) atomicAdd(&hist[(kv[u] >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned acc = 0, digit = 0;
            for (; digit < 256; ++digit) {
                if (acc + hist[digit] >= d) break;
                acc += hist[digit];
            }
            s_prefix = prefix | (digit << shift);
            s_remaining = d - acc;                                          // rank inside the chosen bucket
            s_bucket = hist[digit];                                         // after the last pass: #candidates with key == T
        }
        __syncthreads();
        prefix = s_prefix;
        d = s_remaining;
        mask |= 255u << shift;
        __syncthreads();
    }
    // prefix == T; d == how many candidates with key == T are disabled (the first d by index)
    const unsigned T = prefix;
    const bool all_eq = d == s_bucket;            // every candidate with key == T goes (the usual case: keys are unique)
    for (int base = 0; base < A; base += kSelThreads * kSelUnroll) {
        float lv[kSelUnroll];
        unsigned kv[kSelUnroll];
#pragma unroll
        for (int u = 0; u < kSelUnroll; ++u) {
            const int i = base + u * kSelThreads + threadIdx.x;
            lv[u] = i < A ? lab[i] : -2.f;
            kv[u] = i < A ? key[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < kSelUnroll; ++u)
            if (lv[u] == mine && (kv[u] < T || (all_eq && kv[u] == T))) lab[base + u * kSelThreads + threadIdx.x] = -1.f;
    }
    if (all_eq) return;                                                     // CTA-uniform
    // a tie at the threshold: of the candidates with key == T the first d by index go -- ordered pass
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < A; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const bool eq = i < A && lab[i] == mine && key[i] == T;
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_scan[w];
        const int rank = before + __popc(bal & ((1u << lane) - 1u));
        if (eq && (unsigned)rank < d) lab[i] = -1.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < kSelThreads / 32; ++w) tot += s_scan[w];
            s_base += tot;
        }
        __syncthreads();
        if ((unsigned)s_base >= d) break;                                   // CTA-uniform: all d found
    }
}

// pass 4: regression targets for every inside anchor, loss weights, "unmap" to all A anchors (:125-150)
__global__ void __launch_bounds__(256)
at_finish_kernel(const float4* __restrict__ anchors, int A, int B, const float* __restrict__ gt_left,
                 const float* __restrict__ gt_right, int K, const int* __restrict__ arg_ov,
                 const float* __restrict__ labels, const int* __restrict__ counts, int rpn_batch, int num_fg,
                 float4* __restrict__ tl, float4* __restrict__ tr, float* __restrict__ inside_w,
                 float* __restrict__ outside_w) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    // :136 -- num_examples of the LAST image (the loop variable leaks), after subsampling
    const int fg = counts[(B - 1) * 2], bg = counts[(B - 1) * 2 + 1];
    const int fg_after = min(fg, num_fg);
    const int num_bg = rpn_batch - fg;
    const int bg_after = bg > num_bg ? max(num_bg, 0) : bg;
    const float w = (float)(1.0 / (double)(fg_after + bg_after));
    const size_t o = (size_t)b * A + i;
    const int arg = arg_ov[o];
    float4 dl = make_float4(0.f, 0.f, 0.f, 0.f), dr = dl;
    float iw = 0.f, ow = 0.f;
    if (arg >= 0) {
        const float4 a = anchors[i];
        const float* gl = gt_left + ((size_t)b * K + arg) * 5;
        const float* gr = gt_right + ((size_t)b * K + arg) * 5;
        dl = box_delta(a, make_float4(gl[0], gl[1], gl[2], gl[3]));
        dr = box_delta(a, make_float4(gr[0], gr[1], gr[2], gr[3]));
        const float lab = labels[o];
        iw = lab == 1.f ? 1.f : 0.f;
        ow = lab >= 0.f ? w : 0.f;
    }
    tl[o] = dl;
    tr[o] = dr;
    inside_w[o] = iw;
    outside_w[o] = ow;
}

// --------------------------------------------------------------------------------------- proposal targets
constexpr int kPtThreads = 1024;
constexpr int kPtMaxRois = 4096;      // R + K candidates per image (TRAIN.RPN_POST_NMS_TOP_N 2000 + 30)

struct PtParams {
    const float* rois_left;           // [B][R][5]
    const float* rois_right;
    const float* gt_left;             // [B][K][5]
    const float* gt_right;
    const float* gt_dim_orien;        // [B][K][5]
    const float* gt_kpts;             // [B][K][6]
    const unsigned* keys;             // [B][R+K]
    const unsigned* words;            // [B][S]
    int R, K, S, fg_per_image, grid;
    float fg_thresh, bg_hi, bg_lo;
    float bbox_mean[4], bbox_std[4], dim_mean[5], dim_std[5];
    float* out_rois_left;             // [B][S][5]
    float* out_rois_right;
    float* labels;                    // [B][S]
    float* tl;                        // [B][S][4]
    float* tr;
    float* tdim;                      // [B][S][5]
    int* tkpts;                       // [B][S][3]
    float* wkpts;                     // [B][S][3]
    float* inside_w;                  // [B][S][4]
    float* outside_w;
    int* keep_inds;                   // [B][S]
    int* status;                      // [B]: 0 ok, 1 = neither foreground nor background candidates (:267)
};

__device__ __forceinline__ float4 pt_roi(const float* rois, const float* gt, int b, int R, int K, int i) {
    if (i < R) {
        const float* r = rois + ((size_t)b * R + i) * 5;
        return make_float4(r[1], r[2], r[3], r[4]);
    }
    const float* g = gt + ((size_t)b * K + (i - R)) * 5;      // ground-truth boxes join the candidates (:49-52)
    return make_float4(g[0], g[1], g[2], g[3]);
}

// exclusive block scan of one int per thread (kPtThreads threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_scan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const int base = warp ? s_warp[warp - 1] : 0;
    *total = s_warp[31];
    __syncthreads();
    return base + x - v;
}

__global__ void __launch_bounds__(kPtThreads)
proposal_target_kernel(const PtParams p) {
    __shared__ GtSet sl, sr;
    __shared__ unsigned char s_asg_l[kPtMaxRois], s_asg_r[kPtMaxRois], s_flag[kPtMaxRois];   // flag: 1 fg, 2 bg
    __shared__ unsigned short s_fg[kPtMaxRois], s_bg[kPtMaxRois];
    __shared__ int s_warp[32];
    __shared__ int s_keep[1024];
    const int b = blockIdx.x;
    const int N = p.R + p.K, K = p.K, S = p.S;
    load_gt(sl, p.gt_left + (size_t)b * K * 5, K);
    load_gt(sr, p.gt_right + (size_t)b * K * 5, K);
    __syncthreads();
    // overlaps, assignments, foreground / background predicates (:191-229)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float ml, mr;
        int al, ar;
        max_overlap(sl, K, pt_roi(p.rois_left, p.gt_left, b, p.R, K, i), ml, al);
        max_overlap(sr, K, pt_roi(p.rois_right, p.gt_right, b, p.R, K, i), mr, ar);
        const bool fg = ml >= p.fg_thresh && mr >= p.fg_thresh && al == ar;
        const bool bg = (ml < p.bg_hi && ml >= p.bg_lo) || (mr < p.bg_hi && mr >= p.bg_lo);
        s_asg_l[i] = (unsigned char)al;
        s_asg_r[i] = (unsigned char)ar;
        s_flag[i] = (fg ? 1 : 0) | (bg ? 2 : 0);
    }
    __syncthreads();
    // ordered candidate lists (torch.nonzero / np.union1d order = ascending index)
    int fg_num = 0, bg_num = 0;
    for (int start = 0; start < N; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const int f = i < N ? s_flag[i] : 0;
        int tot;
        const int pf = block_scan(f & 1, s_warp, &tot);
        if (f & 1) s_fg[fg_num + pf] = (unsigned short)i;
        fg_num += tot;
        const int pb = block_scan((f >> 1) & 1, s_warp, &tot);
        if (f & 2) s_bg[bg_num + pb] = (unsigned short)i;
        bg_num += tot;
    }
    __syncthreads();
    int n_fg = 0;
    if (fg_num == 0 && bg_num == 0) {
        if (threadIdx.x == 0) p.status[b] = 1;
        for (int s = threadIdx.x; s < S; s += blockDim.x) s_keep[s] = 0;
    } else {
        if (threadIdx.x == 0) p.status[b] = 0;
        const unsigned* key = p.keys + (size_t)b * N;
        const unsigned* word = p.words + (size_t)b * S;
        if (fg_num > 0 && bg_num > 0) {
            n_fg = min(p.fg_per_image, fg_num);
            // fg_inds[perm[:n_fg]]: the n_fg candidates with the smallest (key, index), in that order
            for (int c = threadIdx.x; c < fg_num; c += blockDim.x) {
                const unsigned kc = key[s_fg[c]];
                int rank = 0;
                for (int o = 0; o < fg_num; ++o) {
                    const unsigned ko = key[s_fg[o]];
                    rank += (ko < kc || (ko == kc && o < c)) ? 1 : 0;
                }
                if (rank < n_fg) s_keep[rank] = s_fg[c];
            }
            for (int j = threadIdx.x; j < S - n_fg; j += blockDim.x)
                s_keep[n_fg + j] = s_bg[(unsigned)(((unsigned long long)word[j] * (unsigned)bg_num) >> 32)];
        } else if (fg_num > 0) {
            n_fg = S;
            for (int j = threadIdx.x; j < S; j += blockDim.x)
                s_keep[j] = s_fg[(unsigned)(((unsigned long long)word[j] * (unsigned)fg_num) >> 32)];
        } else {
            n_fg = 0;
            for (int j = threadIdx.x; j < S; j += blockDim.x)
                s_keep[j] = s_bg[(unsigned)(((unsigned long long)word[j] * (unsigned)bg_num) >> 32)];
        }
    }
    __syncthreads();
    // gather + targets (:269-333), one output row per thread
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const size_t o = (size_t)b * S + s;
        const int i = s_keep[s];
        const bool ok = !(fg_num == 0 && bg_num == 0);
        const int al = s_asg_l[i], ar = s_asg_r[i];
        const float4 rl = pt_roi(p.rois_left, p.gt_left, b, p.R, K, i);
        const float4 rr = pt_roi(p.rois_right, p.gt_right, b, p.R, K, i);
        const float* gl = p.gt_left + ((size_t)b * K + al) * 5;
        const float* gr = p.gt_right + ((size_t)b * K + ar) * 5;
        float lab = (ok && s < n_fg) ? gl[4] : 0.f;                     // :274-277
        p.keep_inds[o] = ok ? i : -1;
        p.labels[o] = lab;
        float* ol = p.out_rois_left + o * 5;
        float* orr = p.out_rois_right + o * 5;
        ol[0] = (float)b; ol[1] = ok ? rl.x : 0.f; ol[2] = ok ? rl.y : 0.f; ol[3] = ok ? rl.z : 0.f; ol[4] = ok ? rl.w : 0.f;
        orr[0] = (float)b; orr[1] = ok ? rr.x : 0.f; orr[2] = ok ? rr.y : 0.f; orr[3] = ok ? rr.z : 0.f; orr[4] = ok ? rr.w : 0.f;
        const bool pos = lab > 0.f;
        float4 dl = make_float4(0.f, 0.f, 0.f, 0.f), dr = dl;
        if (pos) {
            dl = box_delta(rl, make_float4(gl[0], gl[1], gl[2], gl[3]));
            dr = box_delta(rr, make_float4(gr[0], gr[1], gr[2], gr[3]));
            dl.x = __fdiv_rn(__fsub_rn(dl.x, p.bbox_mean[0]), p.bbox_std[0]);
            dl.y = __fdiv_rn(__fsub_rn(dl.y, p.bbox_mean[1]), p.bbox_std[1]);
            dl.z = __fdiv_rn(__fsub_rn(dl.z, p.bbox_mean[2]), p.bbox_std[2]);
            dl.w = __fdiv_rn(__fsub_rn(dl.w, p.bbox_mean[3]), p.bbox_std[3]);
            dr.x = __fdiv_rn(__fsub_rn(dr.x, p.bbox_mean[0]), p.bbox_std[0]);
            dr.y = __fdiv_rn(__fsub_rn(dr.y, p.bbox_mean[1]), p.bbox_std[1]);
            dr.z = __fdiv_rn(__fsub_rn(dr.z, p.bbox_mean[2]), p.bbox_std[2]);
            dr.w = __fdiv_rn(__fsub_rn(dr.w, p.bbox_mean[3]), p.bbox_std[3]);
        }
        reinterpret_cast<float4*>(p.tl)[o] = dl;
        reinterpret_cast<float4*>(p.tr)[o] = dr;
        const float wv = pos ? 1.f : 0.f;
        reinterpret_cast<float4*>(p.inside_w)[o] = make_float4(wv, wv, wv, wv);
        reinterpret_cast<float4*>(p.outside_w)[o] = make_float4(wv, wv, wv, wv);
        const float* gd = p.gt_dim_orien + ((size_t)b * K + al) * 5;
#pragma unroll
        for (int c = 0; c < 5; ++c)
            p.tdim[o * 5 + c] = pos ? __fdiv_rn(__fsub_rn(gd[c], p.dim_mean[c]), p.dim_std[c]) : 0.f;
        // keypoint / border classes (:158-182), kept for class 1 only (:128)
        int tk[3] = {0, 0, 0};
        float wk[3] = {0.f, 0.f, 0.f};
        if (ok && lab == 1.f) {
            const float* gk = p.gt_kpts + ((size_t)b * K + al) * 6;
            const float width = box_w(rl.x, rl.z);
            const float gsz = (float)p.grid;
            float t[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float v = rintf(__fdiv_rn(__fmul_rn(__fsub_rn(gk[c], rl.x), gsz), width));
                if (v < 0.f) v = -225.f;
                if (v > gsz - 1.f) v = -225.f;
                t[c] = v;
            }
            float pos_v = t[0];
            int typ = 0;
#pragma unroll
            for (int c = 1; c < 4; ++c)
                if (t[c] > pos_v) { pos_v = t[c]; typ = c; }
            const float v3[3] = {__fadd_rn(__fmul_rn((float)typ, gsz), pos_v), t[4], t[5]};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                wk[c] = v3[c] < 0.f ? 0.f : 1.f;
                tk[c] = v3[c] < 0.f ? 0 : (int)v3[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p.tkpts[o * 3 + c] = tk[c];
            p.wkpts[o * 3 + c] = wk[c];
        }
    }
}

}  // namespace

extern "C" size_t sb_anchor_targets_workspace(int B, int A) {
    if (B < 1 || A < 1) return 0;
    // max_ov [B][A] f32, arg_ov [B][A] i32, gt_max [B][kMaxGt] u32, counts [B][2] i32
    return (size_t)B * A * 8 + (size_t)B * kMaxGt * 4 + (size_t)B * 2 * 4 + 256;
}

extern "C" int sb_anchor_targets(const float* anchors, int A, const float* gt_left, const float* gt_right,
                                 const float* gt_merge, int B, int K, int im_h, int im_w, const unsigned* keys,
                                 float neg_overlap, float pos_overlap, int rpn_batchsize, int num_fg, void* workspace,
                                 size_t workspace_bytes, float* labels, float* targets_left, float* targets_right,
                                 float* inside_w, float* outside_w, sb_stream_t stream) {
    if (!anchors || !gt_left || !gt_right || !gt_merge || !keys || !workspace || !labels || !targets_left ||
        !targets_right || !inside_w || !outside_w || A < 1 || B < 1 || K < 1 || K > kMaxGt || rpn_batchsize < 1 ||
        num_fg < 0 || workspace_bytes < sb_anchor_targets_workspace(B, A))
        return SB_EINVAL;
    if ((reinterpret_cast<uintptr_t>(anchors) | reinterpret_cast<uintptr_t>(targets_left) |
         reinterpret_cast<uintptr_t>(targets_right) | reinterpret_cast<uintptr_t>(workspace)) & 15)
        return SB_EINVAL;
    cudaStream_t st = sb_cs(stream);
    float* max_ov = static_cast<float*>(workspace);
    int* arg_ov = reinterpret_cast<int*>(max_ov + (size_t)B * A);
    unsigned* gt_max = reinterpret_cast<unsigned*>(arg_ov + (size_t)B * A);
    int* counts = reinterpret_cast<int*>(gt_max + (size_t)B * kMaxGt);
    cudaError_t e = cudaMemsetAsync(gt_max, 0, (size_t)B * kMaxGt * 4 + (size_t)B * 2 * 4, st);
    if (e != cudaSuccess) return (int)e;
    const float4* a4 = reinterpret_cast<const float4*>(anchors);
    const dim3 grid(sb_div_up(A, 256), B);
    if (K <= 32)
        at_overlap_kernel<32><<<grid, 256, 0, st>>>(a4, A, gt_merge, K, (float)im_w, (float)im_h, max_ov, arg_ov, gt_max);
    else
        at_overlap_kernel<kMaxGt><<<grid, 256, 0, st>>>(a4, A, gt_merge, K, (float)im_w, (float)im_h, max_ov, arg_ov, gt_max);
    SB_LAUNCHED();
    at_label_kernel<<<grid, 256, 0, st>>>(a4, A, gt_merge, K, max_ov, arg_ov, gt_max, neg_overlap, pos_overlap, labels,
                                          counts);
    SB_LAUNCHED();
    at_sample_kernel<<<dim3(B, 2), kSelThreads, 0, st>>>(labels, keys, A, counts, rpn_batchsize, num_fg);
    SB_LAUNCHED();
    at_finish_kernel<<<grid, 256, 0, st>>>(a4, A, B, gt_left, gt_right, K, arg_ov, labels, counts, rpn_batchsize, num_fg,
                                           reinterpret_cast<float4*>(targets_left),
                                           reinterpret_cast<float4*>(targets_right), inside_w, outside_w);
    SB_LAUNCHED();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

extern "C" int sb_proposal_targets(const float* rois_left, const float* rois_right, int B, int R, const float* gt_left,
                                   const float* gt_right, const float* gt_dim_orien, const float* gt_kpts, int K,
                                   const unsigned* keys, const unsigned* words, const sb_proposal_target_cfg* cfg,
                                   float* out_rois_left, float* out_rois_right, float* labels, float* bbox_targets_left,
                                   float* bbox_targets_right, float* dim_orien_targets, int* kpts_targets,
                                   float* kpts_weight, float* inside_w, float* outside_w, int* keep_inds, int* status,
                                   sb_stream_t stream) {
    if (!rois_left || !rois_right || !gt_left || !gt_right || !gt_dim_orien || !gt_kpts || !keys || !words || !cfg ||
        !out_rois_left || !out_rois_right || !labels || !bbox_targets_left || !bbox_targets_right ||
        !dim_orien_targets || !kpts_targets || !kpts_weight || !inside_w || !outside_w || !keep_inds || !status)
        return SB_EINVAL;
    if (B < 1 || R < 0 || K < 1 || K > kMaxGt || R + K > kPtMaxRois || cfg->rois_per_image < 1 ||
        cfg->rois_per_image > 1024 || cfg->fg_rois_per_image < 0 || cfg->fg_rois_per_image > cfg->rois_per_image ||
        cfg->kpts_grid < 1)
        return SB_EINVAL;
    if ((reinterpret_cast<uintptr_t>(bbox_targets_left) | reinterpret_cast<uintptr_t>(bbox_targets_right) |
         reinterpret_cast<uintptr_t>(inside_w) | reinterpret_cast<uintptr_t>(outside_w)) & 15)
        return SB_EINVAL;
    PtParams p;
    p.rois_left = rois_left; p.rois_right = rois_right; p.gt_left = gt_left; p.gt_right = gt_right;
    p.gt_dim_orien = gt_dim_orien; p.gt_kpts = gt_kpts; p.keys = keys; p.words = words;
    p.R = R; p.K = K; p.S = cfg->rois_per_image; p.fg_per_image = cfg->fg_rois_per_image; p.grid = cfg->kpts_grid;
    p.fg_thresh = cfg->fg_thresh; p.bg_hi = cfg->bg_thresh_hi; p.bg_lo = cfg->bg_thresh_lo;
    for (int i = 0; i < 4; ++i) { p.bbox_mean[i] = cfg->bbox_means[i]; p.bbox_std[i] = cfg->bbox_stds[i]; }
    for (int i = 0; i < 5; ++i) { p.dim_mean[i] = cfg->dim_means[i]; p.dim_std[i] = cfg->dim_stds[i]; }
    p.out_rois_left = out_rois_left; p.out_rois_right = out_rois_right; p.labels = labels; p.tl = bbox_targets_left;
    p.tr = bbox_targets_right; p.tdim = dim_orien_targets; p.tkpts = kpts_targets; p.wkpts = kpts_weight;
    p.inside_w = inside_w; p.outside_w = outside_w; p.keep_inds = keep_inds; p.status = status;
    proposal_target_kernel<<<B, kPtThreads, 0, sb_cs(stream)>>>(p);
    SB_LAUNCHED();
    SB_CHECK_LAUNCH();
    return SB_OK;
}
