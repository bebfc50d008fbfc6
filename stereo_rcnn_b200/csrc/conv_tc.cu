// conv_tc.cu -- implicit-GEMM convolution on tcgen05 tensor cores (sm_100a only).
//
// The dense contractions of the Stereo R-CNN forward -- every 1x1 / 3x3 convolution of the
// ResNet-101 trunk, FPN, RPN, the 25088->2048 box-head FC, the keypoint-head convs and the
// 2x2 deconv (as four 1x1s) -- run through this one kernel:
//
//   D[M = output pixels, N = Cout] = sum over taps (r,s) and Cin chunks of
//        A[(pixel shifted by tap), Cin chunk] * W[Cout, tap, Cin chunk]^T
//
//   * operands are fp32 in HBM (NHWC activations, [Cout][kh][kw][Cin] weights) and are fed to
//     `tcgen05.mma.kind::tf32` unchanged (the tensor core reads the top 19 bits); accumulation
//     is fp32 in TMEM.  The exact-fp32 SIMT kernel (conv_simt.cu) is the yardstick.
//   * TMA moves every tile: weights through a 2-D map, activations through a 2-D map (1x1:
//     pixels are rows of a flat [M, Cin] matrix) or a 4-D [C, W, H, N] map (3x3: an 8x16 pixel
//     patch per tile; the tap shift is a coordinate offset and the hardware zero-fills the
//     padding halo, so im2col never exists in memory).  128-byte swizzle, 1024-byte aligned
//     stages, one 128-byte row (= 32 fp32 = BLOCK_K) per pixel per stage.
//   * warp-specialised persistent CTAs (one per SM): warp 0 = TMA producer, warp 1 = MMA
//     issuer (one elected lane) + TMEM allocator, warps 2..9 = epilogue (two warps per TMEM lane
//     quarter, each taking half of the tile's columns).  Three mbarrier
//     pipelines: smem full/empty (kStages deep), TMEM full/empty (two accumulator stages, so
//     the epilogue of tile i overlaps the main loop of tile i+1).
//   * fused epilogue straight out of TMEM (`tcgen05.ld.32x32b.x32`): folded frozen-BN
//     scale/shift or bias, residual add, FPN bilinear (align_corners) upsample-add, ReLU,
//     strided / channel-offset stores (RPN L/R concat, deconv scatter).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;   // fp32 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;     // tf32
constexpr int TW = 16, TH = 8;  // spatial patch of a 3x3 tile (TW*TH == BLOCK_M)
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + 32 * kNumEpiWarps;
constexpr int kMaxCout = 2048;   // scale/shift staged in shared memory

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major, 1) | SBO>>4 [32,46) = 8 rows * 128 B
// | version=1 [46,48) | layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=TF32 (2<<7),
// B=TF32 (2<<10), A/B K-major (bits 15/16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ kernel
struct TcParams {
    sb_conv_desc d;
    int patch;            // 0: flat [M,Cin] rows (1x1), 1: 8x16 spatial patches (3x3 / pad 1)
    int tiles_w, tiles_h; // patch mode
    int num_m_tiles, num_n_tiles;
    int kblocks_per_tap;  // Cin / 32
    int num_k_blocks;     // taps * kblocks_per_tap
    long long M;
};

template <int BLOCK_N, int kStages>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const TcParams p) {
    constexpr uint32_t kABytes = BLOCK_M * BLOCK_K * 4, kBBytes = BLOCK_N * BLOCK_K * 4;
    constexpr uint32_t kStageBytes = kABytes + kBBytes;
    constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* tfull = empty + kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);   // [Cout rounded up to BLOCK_N]
    float* s_shift = s_scale + kMaxCout;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const sb_conv_desc& d = p.d;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kNumEpiWarps); }
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    {   // folded-BN scale / shift (or bias) for every output channel, once per (persistent) CTA
        const int cpad = p.num_n_tiles * BLOCK_N;
        for (int c = threadIdx.x; c < cpad; c += kNumThreads) {
            s_scale[c] = (d.scale && c < d.Cout) ? d.scale[c] : 1.f;
            s_shift[c] = (d.shift && c < d.Cout) ? d.shift[c] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
                int n_img = 0, h0 = 0, w0 = 0;
                if (p.patch) {
                    const int tw = mt % p.tiles_w;
                    const int t2 = mt / p.tiles_w;
                    h0 = (t2 % p.tiles_h) * TH;
                    n_img = t2 / p.tiles_h;
                    w0 = tw * TW;
                }
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    const int tap = kb / p.kblocks_per_tap;
                    const int c0 = (kb - tap * p.kblocks_per_tap) * BLOCK_K;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    uint8_t* sb = sa + kABytes;
                    mbar_expect_tx(&full[stage], kStageBytes);
                    if (p.patch) {
                        const int r = tap / d.kw, s = tap - r * d.kw;
                        tma_load_4d(&map_a, &full[stage], sa, c0, w0 + s - d.pad, h0 + r - d.pad, n_img);
                    } else {
                        tma_load_2d(&map_a, &full[stage], sa, c0, mt * BLOCK_M);
                    }
                    tma_load_2d(&map_b, &full[stage], sb, tap * d.Cin + c0, nt * BLOCK_N);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(BLOCK_N);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 32 bytes (8 tf32) inside the 128-byte swizzle row: +2 in 16-byte units
                        umma_tf32(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);                       // frees the smem slot when the MMAs retire
                    if (kb == p.num_k_blocks - 1) umma_commit(&tfull[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // warp w may only touch TMEM lanes 32*(w%4)..+31; two warps share each lane quarter and
        // split the tile's columns in halves.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;          // accumulator row == pixel inside the tile
        constexpr int kColsPerWarp = BLOCK_N / 2 >= 32 ? BLOCK_N / 2 : 32;
        const int col_begin = half * kColsPerWarp;
        const bool has_cols = col_begin < BLOCK_N;
        int acc = 0;
        uint32_t acc_phase = 0;
        float rh = 0.f, rw = 0.f;
        if (d.up_src) {
            rh = d.Ho > 1 ? __fdiv_rn((float)(d.UH - 1), (float)(d.Ho - 1)) : 0.f;
            rw = d.Wo > 1 ? __fdiv_rn((float)(d.UW - 1), (float)(d.Wo - 1)) : 0.f;
        }
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
            int n_img, ho, wo;
            bool valid;
            if (p.patch) {
                const int tw = mt % p.tiles_w;
                const int t2 = mt / p.tiles_w;
                ho = (t2 % p.tiles_h) * TH + row / TW;
                wo = tw * TW + row % TW;
                n_img = t2 / p.tiles_h;
                valid = ho < d.Ho && wo < d.Wo;
            } else {
                const long long m = (long long)mt * BLOCK_M + row;
                valid = m < p.M;
                const long long mm = valid ? m : 0;
                wo = (int)(mm % d.Wo);
                const long long t2 = mm / d.Wo;
                ho = (int)(t2 % d.Ho);
                n_img = (int)(t2 / d.Ho);
            }
            float* orow = d.out + (long long)n_img * d.out_n_stride + (long long)ho * d.out_h_stride +
                          (long long)wo * d.out_w_stride + d.out_coff;
            const float* rrow = d.residual
                                    ? d.residual + ((long long)(n_img * d.Ho + ho) * d.Wo + wo) * d.res_ld
                                    : nullptr;
            const float *u00 = nullptr, *u01 = nullptr, *u10 = nullptr, *u11 = nullptr;
            float ly0 = 0.f, ly1 = 0.f, lx0 = 0.f, lx1 = 0.f;
            if (d.up_src) {
                const float sy = __fmul_rn(rh, (float)ho), sx = __fmul_rn(rw, (float)wo);
                const int y1 = (int)sy, x1 = (int)sx;
                const int yp = y1 < d.UH - 1 ? 1 : 0, xp = x1 < d.UW - 1 ? 1 : 0;
                ly1 = sy - (float)y1; ly0 = 1.f - ly1; lx1 = sx - (float)x1; lx0 = 1.f - lx1;
                u00 = d.up_src + (((long long)n_img * d.UH + y1) * d.UW + x1) * d.Cout;
                u01 = u00 + (long long)xp * d.Cout;
                u10 = u00 + (long long)yp * d.UW * d.Cout;
                u11 = u10 + (long long)xp * d.Cout;
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (has_cols) {
#pragma unroll 1
                for (int cc = col_begin; cc < col_begin + kColsPerWarp; cc += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + cc), v);
                    const int cbase = nt * BLOCK_N + cc;
                    const bool chunk_live = valid && cbase < d.Cout;
                    const bool full = cbase + 31 < d.Cout;
                    // issue the residual loads before waiting on TMEM
                    float4 r4[8];
                    if (chunk_live && rrow && full) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) r4[j] = *reinterpret_cast<const float4*>(rrow + cbase + 4 * j);
                    }
                    tmem_ld_wait();
                    if (chunk_live && full) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = cbase + 4 * j;
                            const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
                            const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
                            float o[4];
                            o[0] = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 0]), sc.x), sh.x);
                            o[1] = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 1]), sc.y), sh.y);
                            o[2] = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 2]), sc.z), sh.z);
                            o[3] = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 3]), sc.w), sh.w);
                            if (rrow) {
                                float4 r = r4[j];
                                if (d.res_biased) {
                                    r.x = sb_unbias_tf32(r.x); r.y = sb_unbias_tf32(r.y);
                                    r.z = sb_unbias_tf32(r.z); r.w = sb_unbias_tf32(r.w);
                                }
                                o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
                            }
                            if (u00) {
                                const float4 a = *reinterpret_cast<const float4*>(u00 + c);
                                const float4 b = *reinterpret_cast<const float4*>(u01 + c);
                                const float4 g = *reinterpret_cast<const float4*>(u10 + c);
                                const float4 h = *reinterpret_cast<const float4*>(u11 + c);
                                o[0] += ly0 * (lx0 * a.x + lx1 * b.x) + ly1 * (lx0 * g.x + lx1 * h.x);
                                o[1] += ly0 * (lx0 * a.y + lx1 * b.y) + ly1 * (lx0 * g.y + lx1 * h.y);
                                o[2] += ly0 * (lx0 * a.z + lx1 * b.z) + ly1 * (lx0 * g.z + lx1 * h.z);
                                o[3] += ly0 * (lx0 * a.w + lx1 * b.w) + ly1 * (lx0 * g.w + lx1 * h.w);
                            }
                            if (d.relu) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
                            }
                            if (d.out_mode) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) o[e] = sb_store_mode(o[e], d.out_mode);
                            }
                            *reinterpret_cast<float4*>(orow + c) = make_float4(o[0], o[1], o[2], o[3]);
                        }
                    } else if (chunk_live) {
                        // ragged tail of the channel range (Cout not a multiple of 32): scalar path
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int c = cbase + j;
                            if (c >= d.Cout) continue;
                            float x = __fadd_rn(__fmul_rn(__uint_as_float(v[j]), s_scale[c]), s_shift[c]);
                            if (rrow) x += d.res_biased ? sb_unbias_tf32(rrow[c]) : rrow[c];
                            if (u00)
                                x += ly0 * (lx0 * u00[c] + lx1 * u01[c]) + ly1 * (lx0 * u10[c] + lx1 * u11[c]);
                            if (d.relu) x = fmaxf(x, 0.f);
                            orow[c] = sb_store_mode(x, d.out_mode);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims,
                     strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Tile-width choice.  One SM ingests ~64 B/clk from L2, so a K-step costs max(MMA cycles, bytes/64):
//   BN=128: max(256, (16+16) KB / 64) = 512 cycles     BN=256: max(512, (16+32) KB / 64) = 768 cycles
// and the persistent grid runs ceil(tiles/SMs) waves.  Pick the cheaper of the two for Cout >= 256.
int pick_block_n(int cout, long long m_tiles, int num_sms) {
    if (cout <= 32) return 32;
    if (cout <= 64) return 64;
    if (cout < 256) return 128;
    const long long t128 = m_tiles * ((cout + 127) / 128), t256 = m_tiles * ((cout + 255) / 256);
    const long long c128 = ((t128 + num_sms - 1) / num_sms) * 512, c256 = ((t256 + num_sms - 1) / num_sms) * 768;
    return c256 <= c128 ? 256 : 128;
}

template <int BN, int ST>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)ST * (BLOCK_M * BLOCK_K * 4 + BN * BLOCK_K * 4) + 1024 + 256 + 2 * kMaxCout * 4;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = tiles < num_sms ? tiles : num_sms;
    conv_tc_kernel<BN, ST><<<grid, kNumThreads, smem, st>>>(ma, mb, p);
    SB_LAUNCHED();
    SB_CHECK_LAUNCH();
    return SB_OK;
}

}  // namespace

extern "C" int sb_conv2d_tc_supported(const sb_conv_desc* d) {
    if (!d || !d->in || !d->wgt || !d->out) return 0;
    if (d->stride != 1 || d->Cin % BLOCK_K != 0 || d->in_ld % 4 != 0) return 0;
    const bool k1 = d->kh == 1 && d->kw == 1 && d->pad == 0;
    const bool k3 = d->kh == 3 && d->kw == 3 && d->pad == 1;
    if (!k1 && !k3) return 0;
    if (d->Cout > kMaxCout) return 0;
    if ((reinterpret_cast<uintptr_t>(d->in) & 15) || (reinterpret_cast<uintptr_t>(d->wgt) & 15) ||
        (reinterpret_cast<uintptr_t>(d->out) & 15))
        return 0;
    if ((d->out_coff & 3) || (d->out_n_stride & 3) || (d->out_h_stride & 3) || (d->out_w_stride & 3)) return 0;
    if (d->residual && ((d->res_ld & 3) || (reinterpret_cast<uintptr_t>(d->residual) & 15))) return 0;
    if (d->up_src && ((d->Cout & 3) || (reinterpret_cast<uintptr_t>(d->up_src) & 15))) return 0;
    if (d->Ho != d->H || d->Wo != d->W) return 0;
    return 1;
}

extern "C" int sb_conv2d_tc(const sb_conv_desc* d, sb_stream_t stream) {
    if (!sb_conv2d_tc_supported(d)) return SB_EINVAL;
    TcParams p;
    p.d = *d;
    p.patch = (d->kh == 3) ? 1 : 0;
    p.M = (long long)d->N * d->Ho * d->Wo;
    if (p.M == 0) return SB_OK;
    p.kblocks_per_tap = d->Cin / BLOCK_K;
    p.num_k_blocks = d->kh * d->kw * p.kblocks_per_tap;
    p.tiles_w = (d->W + TW - 1) / TW;
    p.tiles_h = (d->H + TH - 1) / TH;
    p.num_m_tiles = p.patch ? d->N * p.tiles_h * p.tiles_w : (int)((p.M + BLOCK_M - 1) / BLOCK_M);
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    int BN = pick_block_n(d->Cout, p.num_m_tiles, sms);
    if (const char* e = getenv("SB_TC_BLOCK_N")) { int v = atoi(e); if ((v == 128 || v == 256) && d->Cout >= 256) BN = v; }
    p.num_n_tiles = (d->Cout + BN - 1) / BN;
    CUtensorMap ma, mb;
    if (p.patch) {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->in_ld * 4, (cuuint64_t)d->W * d->in_ld * 4,
                                 (cuuint64_t)d->H * d->W * d->in_ld * 4};
        cuuint32_t box[4] = {BLOCK_K, TW, TH, 1};
        if (!make_map(&ma, d->in, 4, dims, strides, box)) return SB_EINVAL;
    } else {
        cuuint64_t dims[2] = {(cuuint64_t)d->Cin, (cuuint64_t)p.M};
        cuuint64_t strides[1] = {(cuuint64_t)d->in_ld * 4};
        cuuint32_t box[2] = {BLOCK_K, BLOCK_M};
        if (!make_map(&ma, d->in, 2, dims, strides, box)) return SB_EINVAL;
    }
    {
        const cuuint64_t ktot = (cuuint64_t)d->kh * d->kw * d->Cin;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)d->Cout};
        cuuint64_t strides[1] = {ktot * 4};
        cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)BN};
        if (!make_map(&mb, d->wgt, 2, dims, strides, box)) return SB_EINVAL;
    }
    cudaStream_t st = sb_cs(stream);
    switch (BN) {
        case 32: return launch<32, 8>(ma, mb, p, st);
        case 64: return launch<64, 8>(ma, mb, p, st);
        case 256: return launch<256, 4>(ma, mb, p, st);
        default: return launch<128, 5>(ma, mb, p, st);
    }
}
