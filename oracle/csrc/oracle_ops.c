/*
 * oracle_ops.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Plain-C restatement of the reference's arithmetic for the memory-bound
 * operators of the Stereo R-CNN hot path.  Nothing in the product path may
 * link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference checkout).  Compile with -ffp-contract=off so that every
 * fp32 operation below rounds exactly once, as written; explicit fmaf() is
 * used only where stated.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* deterministic expf shared bit-for-bit with the CUDA kernels         */
/* (csrc/common.cuh: sb_expf).  <= 1 ulp from a correctly rounded exp  */
/* on the RPN delta range; stands in for torch.exp at                  */
/* lib/model/rpn/bbox_transform.py:93-94.                              */
/* ------------------------------------------------------------------ */
float sb_expf(float x) {
    if (x > 88.72283f) return INFINITY;
    if (x < -103.9f) return 0.0f;
    const float log2e = 1.44269504088896341f;
    const float ln2_hi = 0.693145751953125f;       /* 12 trailing zero bits */
    const float ln2_lo = 1.42860682030941723212e-6f;
    float n = rintf(x * log2e);
    float r = fmaf(n, -ln2_hi, x);
    r = fmaf(n, -ln2_lo, r);
    /* degree-7 Taylor (truncation 5e-9 on |r| <= 0.3466) */
    float p = 1.0f / 5040.0f;
    p = fmaf(p, r, 1.0f / 720.0f);
    p = fmaf(p, r, 1.0f / 120.0f);
    p = fmaf(p, r, 1.0f / 24.0f);
    p = fmaf(p, r, 1.0f / 6.0f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int ni = (int)n;
    /* scale by 2^ni in two steps so that subnormal / overflow edges behave */
    int n1 = ni / 2, n2 = ni - n1;
    union { uint32_t u; float f; } s1, s2;
    s1.u = (uint32_t)(n1 + 127) << 23;
    s2.u = (uint32_t)(n2 + 127) << 23;
    return (p * s1.f) * s2.f;
}

void sb_expf_array(const float* x, float* y, int n) {
    for (int i = 0; i < n; ++i) y[i] = sb_expf(x[i]);
}

/* ------------------------------------------------------------------ */
/* NMS: lib/model/nms/src/nms_cuda_kernel.cu:31-39 (devIoU),           */
/* :41-85 (mask), :132-144 (greedy host scan).                         */
/* dets: n x 5 [x1,y1,x2,y2,score], pre-sorted by score desc.          */
/* ------------------------------------------------------------------ */
static inline float dev_iou(const float* a, const float* b) {
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float width = fmaxf(right - left + 1.0f, 0.f);
    float height = fmaxf(bottom - top + 1.0f, 0.f);
    float interS = width * height;
    float Sa = (a[2] - a[0] + 1.0f) * (a[3] - a[1] + 1.0f);
    float Sb = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
    return interS / (Sa + Sb - interS);
}

/* returns number kept; keep[] receives ascending indices into dets */
int oracle_nms(const float* dets, int n, int stride, float thresh, int* keep) {
    unsigned char* dead = (unsigned char*)calloc((size_t)n + 1, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        const float* a = dets + (size_t)i * stride;
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            if (dev_iou(a, dets + (size_t)j * stride) > thresh) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}

/* full 64-bit suppression mask exactly as nms_kernel writes it (for mask
 * level parity tests): mask[i*col_blocks + cb] bit b set iff
 * IoU(i, cb*64+b) > thresh and (cb*64+b) > i within the diagonal block,
 * all j in off-diagonal blocks (the reference computes the full square). */
void oracle_nms_mask(const float* dets, int n, int stride, float thresh,
                     uint64_t* mask) {
    int cbn = (n + 63) / 64;
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        for (int cb = 0; cb < cbn; ++cb) {
            uint64_t t = 0;
            int start = (i / 64 == cb) ? (i % 64) + 1 : 0;
            int cs = n - cb * 64 < 64 ? n - cb * 64 : 64;
            for (int b = start; b < cs; ++b)
                if (dev_iou(dets + (size_t)i * stride,
                            dets + (size_t)(cb * 64 + b) * stride) > thresh)
                    t |= 1ULL << b;
            mask[(size_t)i * cbn + cb] = t;
        }
    }
}

/* ------------------------------------------------------------------ */
/* RoIAlign forward: lib/model/roi_align/src/roi_align_kernel.cu:15-70 */
/* features NCHW, rois r x 5 [b,x1,y1,x2,y2], out r x C x ah x aw      */
/* (ah,aw are the *lattice* sizes, i.e. RoIAlignAvg passes pooled+1).  */
/* Double-precision sub-expressions follow the `1.` literals of the    */
/* source; the h/w lattice coordinate is one fused multiply-add, which */
/* is what nvcc emits for that line under its default -fmad=true.      */
/* ------------------------------------------------------------------ */
typedef struct {
    float start_w, start_h, bin_w, bin_h;
    int batch;
} roi_geom_t;

static inline roi_geom_t roi_geom(const float* roi, float scale, int ah, int aw) {
    roi_geom_t g;
    g.batch = (int)roi[0];
    float sw = roi[1] * scale, sh = roi[2] * scale;
    float ew = roi[3] * scale, eh = roi[4] * scale;
    float rw = fmaxf((float)((double)(ew - sw) + 1.), 0.f);
    float rh = fmaxf((float)((double)(eh - sh) + 1.), 0.f);
    g.bin_h = (float)((double)rh / ((double)ah - 1.));
    g.bin_w = (float)((double)rw / ((double)aw - 1.));
    g.start_w = sw;
    g.start_h = sh;
    return g;
}

void oracle_roi_align_forward(const float* feat, int N, int C, int H, int W,
                              const float* rois, int R, int ah, int aw,
                              float scale, float* out) {
    (void)N;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < R; ++n) {
        roi_geom_t g = roi_geom(rois + 5 * n, scale, ah, aw);
        for (int ph = 0; ph < ah; ++ph) {
            float h = fmaf((float)ph, g.bin_h, g.start_h);
            int hstart = (int)fminf(floorf(h), (float)(H - 2));
            for (int pw = 0; pw < aw; ++pw) {
                float w = fmaf((float)pw, g.bin_w, g.start_w);
                int wstart = (int)fminf(floorf(w), (float)(W - 2));
                int zero = (h < 0 || h >= H || w < 0 || w >= W);
                float hr = h - (float)hstart, wr = w - (float)wstart;
                for (int c = 0; c < C; ++c) {
                    float* o = out + (((size_t)n * C + c) * ah + ph) * aw + pw;
                    if (zero) { *o = 0.f; continue; }
                    const float* f = feat + (((size_t)g.batch * C + c) * H + hstart) * W + wstart;
                    double v = (double)f[0] * (1. - hr) * (1. - wr)
                             + (double)f[1] * (1. - hr) * wr
                             + (double)f[W] * hr * (1. - wr)
                             + (double)f[W + 1] * hr * wr;
                    *o = (float)v;
                }
            }
        }
    }
}

/* RoIAlign backward: roi_align_kernel.cu:94-143.  The reference scatters
 * with atomicAdd (order undefined); the oracle accumulates in double and
 * rounds once, which every summation order agrees with to ~1 ulp.        */
void oracle_roi_align_backward(const float* top, int N, int C, int H, int W,
                               const float* rois, int R, int ah, int aw,
                               float scale, float* bottom) {
    size_t tot = (size_t)N * C * H * W;
    double* acc = (double*)calloc(tot, sizeof(double));
    for (int n = 0; n < R; ++n) {
        roi_geom_t g = roi_geom(rois + 5 * n, scale, ah, aw);
        for (int ph = 0; ph < ah; ++ph) {
            float h = fmaf((float)ph, g.bin_h, g.start_h);
            int hstart = (int)fminf(floorf(h), (float)(H - 2));
            for (int pw = 0; pw < aw; ++pw) {
                float w = fmaf((float)pw, g.bin_w, g.start_w);
                int wstart = (int)fminf(floorf(w), (float)(W - 2));
                if (h < 0 || h >= H || w < 0 || w >= W) continue;
                float hr = h - (float)hstart, wr = w - (float)wstart;
                for (int c = 0; c < C; ++c) {
                    double t = top[(((size_t)n * C + c) * ah + ph) * aw + pw];
                    double* b = acc + (((size_t)g.batch * C + c) * H + hstart) * W + wstart;
                    b[0] += (double)(float)(t * (1. - hr) * (1 - wr));
                    b[1] += (double)(float)(t * (1. - hr) * wr);
                    b[W] += (double)(float)(t * hr * (1 - wr));
                    b[W + 1] += (double)(float)(t * hr * wr);
                }
            }
        }
    }
    for (size_t i = 0; i < tot; ++i) bottom[i] = (float)acc[i];
    free(acc);
}

/* ------------------------------------------------------------------ */
/* Box decode + clip: lib/model/rpn/bbox_transform.py:79-104,177-185   */
/* (torch evaluates each op separately: no contraction), exp replaced  */
/* by sb_expf (see header).                                            */
/* ------------------------------------------------------------------ */
void oracle_decode_clip(const float* boxes, const float* deltas, int n,
                        float im_h, float im_w, float* out) {
    float xmax = im_w - 1.0f, ymax = im_h - 1.0f;
    for (int i = 0; i < n; ++i) {
        const float* b = boxes + 4 * i;
        const float* d = deltas + 4 * i;
        float w = b[2] - b[0] + 1.0f, h = b[3] - b[1] + 1.0f;
        float cx = b[0] + 0.5f * w, cy = b[1] + 0.5f * h;
        float pcx = d[0] * w + cx, pcy = d[1] * h + cy;
        float pw = sb_expf(d[2]) * w, ph = sb_expf(d[3]) * h;
        float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph;
        float x2 = pcx + 0.5f * pw, y2 = pcy + 0.5f * ph;
        out[4 * i + 0] = fminf(fmaxf(x1, 0.f), xmax);
        out[4 * i + 1] = fminf(fmaxf(y1, 0.f), ymax);
        out[4 * i + 2] = fminf(fmaxf(x2, 0.f), xmax);
        out[4 * i + 3] = fminf(fmaxf(y2, 0.f), ymax);
    }
}

/* ------------------------------------------------------------------ */
/* dense_align: lib/model/dense_align/dense_align.py:13-69,175-300 and */
/* lib/model/dense_align/box_3d.py:12-106.         */
/* ------------------------------------------------------------------ */

/* F.upsample(scale_factor=2, mode='bilinear') with align_corners=True
 * (torch 0.3.0 semantics, dense_align.py:256-257).  src C x H x W planar
 * -> dst C x 2H x 2W planar.                                            */
void oracle_upsample2x(const float* src, int C, int H, int W, float* dst) {
    int OH = 2 * H, OW = 2 * W;
    float rh = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    float rw = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = 0; c < C; ++c)
        for (int y = 0; y < OH; ++y) {
            float sy = rh * (float)y;
            int y1 = (int)sy;
            int yp = (y1 < H - 1) ? 1 : 0;
            float ly1 = sy - (float)y1, ly0 = 1.f - ly1;
            const float* r0 = src + ((size_t)c * H + y1) * W;
            const float* r1 = r0 + (size_t)yp * W;
            float* o = dst + ((size_t)c * OH + y) * OW;
            for (int x = 0; x < OW; ++x) {
                float sx = rw * (float)x;
                int x1 = (int)sx;
                int xp = (x1 < W - 1) ? 1 : 0;
                float lx1 = sx - (float)x1, lx0 = 1.f - lx1;
                o[x] = ly0 * (lx0 * r0[x1] + lx1 * r0[x1 + xp]) +
                       ly1 * (lx0 * r1[x1] + lx1 * r1[x1 + xp]);
            }
        }
}

/* python slice(start, stop, step>0) on a dimension of length size */
static int py_slice(int start, int stop, int step, int size, int* first) {
    if (start < 0) { start += size; if (start < 0) start = 0; }
    if (start > size) start = size;
    if (stop < 0) { stop += size; if (stop < 0) stop = 0; }
    if (stop > size) stop = size;
    *first = start;
    if (stop <= start) return 0;
    return (stop - start + step - 1) / step;
}

typedef struct {
    /* per-RoI geometry (box_3d.py:13-60) */
    float T[3];
    float c, s;            /* cos/sin(theta): double math, rounded to fp32 */
    float pmin[3], pmax[3];/* P_o[4]-eps, P_o[2]+eps (box_3d.py:72-77)     */
    float planes[3][4];    /* the three visible planes in fill order       */
    /* pixel lattice (dense_align.py:39-45) */
    int u0, nu, su, v0, nv, sv;
} da_roi_t;

static void make_plane(const float* p1, const float* p2, const float* p3, float* pl) {
    float a1[3], a2[3];
    for (int k = 0; k < 3; ++k) { a1[k] = p2[k] - p1[k]; a2[k] = p3[k] - p1[k]; }
    float n0 = a1[1] * a2[2] - a1[2] * a2[1];
    float n1 = a1[2] * a2[0] - a1[0] * a2[2];
    float n2 = a1[0] * a2[1] - a1[1] * a2[0];
    pl[0] = n0; pl[1] = n1; pl[2] = n2;
    pl[3] = (-n0 * p1[0] - n1 * p1[1]) - n2 * p1[2];
}

static const int PLANE_VERTS[6][3] = {
    {0, 3, 4}, {2, 3, 6}, {1, 2, 5}, {0, 1, 4}, {0, 1, 2}, {4, 5, 6}};
static const int PLANE_GROUP[8][3] = {
    {0, 3, 4}, {2, 3, 4}, {1, 2, 4}, {0, 1, 4},
    {0, 3, 5}, {2, 3, 5}, {1, 2, 5}, {0, 1, 5}};

/* box (already x s2), border (already x s2), pose: raw.  fh,fw: upsampled size */
static void da_setup(const float* box, const float* border, const float* pose,
                     int fh, int fw, da_roi_t* g) {
    g->T[0] = pose[0]; g->T[1] = pose[1]; g->T[2] = pose[2];
    float w = pose[3], h = pose[4], l = pose[5];
    g->c = (float)cos((double)pose[6]);
    g->s = (float)sin((double)pose[6]);
    float hw = w / 2.0f, hl = l / 2.0f;
    float Po[8][3] = {
        {-hw, 0.f, -hl}, {-hw, 0.f, hl}, {hw, 0.f, hl}, {hw, 0.f, -hl},
        {-hw, -h, -hl},  {-hw, -h, hl},  {hw, -h, hl},  {hw, -h, -hl}};
    float Pc[8][3];
    for (int i = 0; i < 8; ++i) {
        /* torch.mm(R, P_o) + T, R = [[c,0,s],[0,1,0],[-s,0,c]] (box_3d.py:17-33) */
        Pc[i][0] = ((g->c * Po[i][0] + 0.f * Po[i][1]) + g->s * Po[i][2]) + g->T[0];
        Pc[i][1] = ((0.f * Po[i][0] + 1.f * Po[i][1]) + 0.f * Po[i][2]) + g->T[1];
        Pc[i][2] = ((-g->s * Po[i][0] + 0.f * Po[i][1]) + g->c * Po[i][2]) + g->T[2];
    }
    int nearest = 0;
    float best = 100000000.f;
    for (int i = 0; i < 8; ++i) {
        float nn = sqrtf((Pc[i][0] * Pc[i][0] + Pc[i][1] * Pc[i][1]) + Pc[i][2] * Pc[i][2]);
        if (nn < best) { best = nn; nearest = i; }
    }
    for (int k = 0; k < 3; ++k) {
        const int* v = PLANE_VERTS[PLANE_GROUP[nearest][k]];
        make_plane(Pc[v[0]], Pc[v[1]], Pc[v[2]], g->planes[k]);
    }
    g->pmin[0] = -hw - 0.01f; g->pmin[1] = -h - 0.01f; g->pmin[2] = -hl - 0.01f;
    g->pmax[0] = hw + 0.01f;  g->pmax[1] = 0.f + 0.01f; g->pmax[2] = hl + 0.01f;
    /* lattice (dense_align.py:39-45); int() truncates toward zero */
    int su = (int)((border[1] - border[0]) / 56.0f); if (su < 1) su = 1;
    int sv = (int)((box[3] - box[1]) / 56.0f);       if (sv < 1) sv = 1;
    int vs = (int)((box[1] + box[3]) / 2.0f + 0.5f);
    int ve = (int)(box[3] - (box[3] - box[1]) * 0.1f + 0.5f);
    int us = (int)(border[0] + 0.5f);
    int ue = (int)(border[1] + 0.5f);
    g->su = su; g->sv = sv;
    g->nv = py_slice(vs, ve, sv, fh, &g->v0);
    g->nu = py_slice(us, ue, su, fw, &g->u0);
}

/* ray/box test for pixel (u,v): returns 1 and *dz when valid (box_3d.py:62-106) */
static int da_ray(const da_roi_t* g, float u, float v, float cx, float cy, float f,
                  float* dz) {
    float rx = (u - cx) / f, ry = (v - cy) / f;
    float ox = 0.f, oy = 0.f, oz = 0.f, m = 0.f;
    (void)ox; (void)oy;
    for (int k = 0; k < 3; ++k) {
        if (m != 0.f) break; /* only still-invalid pixels are overwritten */
        const float* pl = g->planes[k];
        float t = (rx * pl[0] + ry * pl[1]) + 1.0f * pl[2];
        t = -(1.0f / t) * pl[3];
        float ix = rx * t - g->T[0], iy = ry * t - g->T[1], iz = 1.0f * t - g->T[2];
        float bx = (g->c * ix + 0.f * iy) + (-g->s) * iz;
        float by = (0.f * ix + 1.f * iy) + 0.f * iz;
        float bz = (g->s * ix + 0.f * iy) + g->c * iz;
        int in = (bx >= g->pmin[0]) && (by >= g->pmin[1]) && (bz >= g->pmin[2]) &&
                 (bx <= g->pmax[0]) && (by <= g->pmax[1]) && (bz <= g->pmax[2]);
        oz = iz;
        m = in ? 1.f : 0.f;
    }
    *dz = oz;
    return m != 0.f;
}

/* F.grid_sample(bilinear, padding_mode='border', align_corners=True) at
 * normalised (gx,gy) on a planar C x H x W image; returns 3 channels.
 * Tap/weight arithmetic follows ATen's grid_sampler_2d (unnormalise,
 * clip, nw/ne/sw/se weights).                                          */
static void grid_sample3(const float* im, int H, int W, float gx, float gy, float* out) {
    float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    float x0 = floorf(ix), y0 = floorf(iy);
    float x1 = x0 + 1.f, y1 = y0 + 1.f;
    float nw = (x1 - ix) * (y1 - iy), ne = (ix - x0) * (y1 - iy);
    float sw = (x1 - ix) * (iy - y0), se = (ix - x0) * (iy - y0);
    int xi0 = (int)x0, yi0 = (int)y0, xi1 = xi0 + 1, yi1 = yi0 + 1;
    int okx1 = xi1 <= W - 1, oky1 = yi1 <= H - 1;
    for (int c = 0; c < 3; ++c) {
        const float* p = im + (size_t)c * H * W;
        float v = p[(size_t)yi0 * W + xi0] * nw;
        if (okx1) v += p[(size_t)yi0 * W + xi1] * ne;
        if (oky1) v += p[(size_t)yi1 * W + xi0] * sw;
        if (okx1 && oky1) v += p[(size_t)yi1 * W + xi1] * se;
        out[c] = v;
    }
}

/*
 * Whole align_parallel (dense_align.py:240-300).
 *   im_left/right : 3 x H x W planar fp32 (network-scale images)
 *   calib         : {P2[0,0], P2[0,2], P2[1,2], P2[0,3]-P3[0,3]} as double
 *   scale         : im_info[0,2] (python float of an fp32 value)
 *   box_left D x 4, keypoints D x 5, poses D x 7 (original-image units)
 * outputs: status[D], best_dis[D]; optional diagnostics npix[D] (valid pixel
 * count), cost_coarse[D*50], cost_fine[D*20] (double-accumulated SAD),
 * idx[D*2] (argmin indices); pass NULL to skip.
 */
void oracle_dense_align(const float* im_left, const float* im_right, int H, int W,
                        const double* calib, double scale,
                        const float* box_left, const float* keypoints,
                        const float* poses, int D,
                        float* status, float* best_dis,
                        int* npix, double* cost_coarse, double* cost_fine, int* idx) {
    double s2 = scale * 2.0;
    int FH = 2 * H, FW = 2 * W;
    float* upL = (float*)malloc(sizeof(float) * 3 * (size_t)FH * FW);
    float* upR = (float*)malloc(sizeof(float) * 3 * (size_t)FH * FW);
    oracle_upsample2x(im_left, 3, H, W, upL);
    oracle_upsample2x(im_right, 3, H, W, upR);

    double fd = calib[0] * s2;
    double bld = calib[3] * s2 / fd;
    float s2f = (float)s2;
    float f32 = (float)fd, bl32 = (float)bld, fb32 = (float)(fd * bld);
    float cx32 = (float)(calib[1] * s2), cy32 = (float)(calib[2] * s2);
    float fw2 = (float)(((double)FW - 1.0) / 2.0), fh2 = (float)(((double)FH - 1.0) / 2.0);

    int any_valid = 0;
    float* dis_init = (float*)malloc(sizeof(float) * (D > 0 ? D : 1));

#pragma omp parallel for schedule(dynamic, 1) reduction(| : any_valid)
    for (int i = 0; i < D; ++i) {
        float box[4], border[2];
        for (int k = 0; k < 4; ++k) box[k] = box_left[4 * i + k] * s2f;
        border[0] = keypoints[5 * i + 3] * s2f;
        border[1] = keypoints[5 * i + 4] * s2f;
        const float* pose = poses + 7 * i;
        dis_init[i] = fb32 / pose[2];
        da_roi_t g;
        da_setup(box, border, pose, FH, FW, &g);
        int cap = g.nu * g.nv;
        float* pu = (float*)malloc(sizeof(float) * (cap > 0 ? cap : 1) * 6);
        float *pv = pu + cap, *pz = pv + cap, *pl = pz + cap; /* pl: 3 floats / px */
        int P = 0;
        for (int a = 0; a < g.nv; ++a)
            for (int b = 0; b < g.nu; ++b) {
                float u = (float)(g.u0 + b * g.su), v = (float)(g.v0 + a * g.sv), dz;
                if (da_ray(&g, u, v, cx32, cy32, f32, &dz)) {
                    pu[P] = u; pv[P] = v; pz[P] = dz;
                    grid_sample3(upL, FH, FW, (u - fw2) / fw2, (v - fh2) / fh2, pl + 3 * P);
                    ++P;
                }
            }
        if (npix) npix[i] = P;
        status[i] = P > 0 ? 1.f : 0.f;
        if (P > 0) any_valid = 1;

        /* coarse: depth_enum[i] = 1/dis_init * f * bl - 12.5 + 0.5 i, clamp >= 1.5 */
        float z0 = ((1.0f / dis_init[i]) * f32) * bl32;
        float best_depth = 0.f;
        for (int stage = 0; stage < 2; ++stage) {
            int nh = stage == 0 ? 50 : 20;
            double best_cost = 0; int best_i = -1; float best_d = 0.f;
            for (int hi = 0; hi < nh; ++hi) {
                float depth;
                if (stage == 0) {
                    depth = (z0 - 12.5f) + (float)(0.5 * hi);
                    if (depth < 1.5f) depth = 1.5f;
                } else {
                    /* best - tune_num*tune_interval/2 + tune_interval*i (dense_align.py:291-294) */
                    depth = (best_depth - 0.5f) + (float)(0.05 * hi);
                }
                float dis = (1.0f / depth) * fb32;
                float rdis = 1.0f / dis;
                double cost = 0.0;
                for (int p = 0; p < P; ++p) {
                    float d = 1.0f / (pz[p] / fb32 + rdis);
                    float gx = ((pu[p] - d) - fw2) / fw2;
                    float gy = (pv[p] - fh2) / fh2;
                    float r[3];
                    grid_sample3(upR, FH, FW, gx, gy, r);
                    for (int c = 0; c < 3; ++c) cost += (double)fabsf(pl[3 * p + c] - r[c]);
                }
                if (stage == 0 && cost_coarse) cost_coarse[(size_t)i * 50 + hi] = cost;
                if (stage == 1 && cost_fine) cost_fine[(size_t)i * 20 + hi] = cost;
                if (best_i < 0 || cost < best_cost) { best_cost = cost; best_i = hi; best_d = depth; }
            }
            best_depth = best_d;
            if (idx) idx[2 * i + stage] = best_i;
        }
        best_dis[i] = fb32 / (best_depth * s2f) + 0.5f;
        free(pu);
    }
    if (!any_valid) { /* dense_align.py:272-273 early return */
        for (int i = 0; i < D; ++i) { status[i] = 0.f; best_dis[i] = dis_init[i]; }
    }
    free(dis_init);
    free(upL);
    free(upR);
}

/* sample() alone (dense_align.py:13-69): fills uvz (D x maxp x 3) row-major
 * compacted valid pixels; returns per-RoI counts.  maxp is a capacity.   */
void oracle_dense_sample(int H, int W, const double* calib, double scale,
                         const float* box_left, const float* keypoints,
                         const float* poses, int D, int maxp,
                         float* uvz, int* npix) {
    double s2 = scale * 2.0;
    int FH = 2 * H, FW = 2 * W;
    float s2f = (float)s2;
    float f32 = (float)(calib[0] * s2);
    float cx32 = (float)(calib[1] * s2), cy32 = (float)(calib[2] * s2);
    for (int i = 0; i < D; ++i) {
        float box[4], border[2];
        for (int k = 0; k < 4; ++k) box[k] = box_left[4 * i + k] * s2f;
        border[0] = keypoints[5 * i + 3] * s2f;
        border[1] = keypoints[5 * i + 4] * s2f;
        da_roi_t g;
        da_setup(box, border, poses + 7 * i, FH, FW, &g);
        int P = 0;
        for (int a = 0; a < g.nv; ++a)
            for (int b = 0; b < g.nu; ++b) {
                float u = (float)(g.u0 + b * g.su), v = (float)(g.v0 + a * g.sv), dz;
                if (da_ray(&g, u, v, cx32, cy32, f32, &dz)) {
                    if (P < maxp) {
                        float* o = uvz + ((size_t)i * maxp + P) * 3;
                        o[0] = u; o[1] = v; o[2] = dz;
                    }
                    ++P;
                }
            }
        npix[i] = P;
    }
}
