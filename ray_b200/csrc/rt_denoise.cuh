// rt_denoise.cuh -- RendererBase::DenoiseImage(const RegionContext &): joint non-local-means filter
// (first half).
//
// Behavioural spec: reference internal/RendererCPU.h:661-787 (variance blur, required-samples update, invert + tonemap) and
// internal/DenoiseRef.cpp:9-93 (JointNLMFilter<7, 3> with base colour and depth-normals as features).  Every pixel's
// arithmetic -- including the reference's (x - i + 1) tap positions of the 9-tap blur, the 4-lane accumulation order
// and the libm expf -- is reproduced, so the filtered linear image is bit-identical; the tonemapped plane differs only
// through powf (as in k_resolve).
//
// Kernel shape: the filter reads a 7x7 window of 3x3 patches: per output pixel 49 x 9 x 2 float4 from two planes, all
// inside a radius-4 halo.  A 32x8 block stages its (40 x 16) halo tile of both planes in shared memory once (20 KB),
// so the planes are read from L2/HBM ~1.25x instead of 882x; the feature planes (49 x 2 float4 per pixel) come through L1.
#pragma once

#include "rt_kernels.cuh"

namespace rt {

constexpr int kNlmExt = 8;      // EXT_RADIUS
constexpr int kNlmWindow = 3;   // (7 - 1) / 2
constexpr int kNlmPatch = 1;    // (3 - 1) / 2

struct NlmParams {
    FrameBufs fb;
    int rx, ry, rw, rh; // region
    int ex, ey, ew, eh; // region grown by kNlmExt
    float4 *temp_final, *var_h, *var_f; // ew * eh scratch planes
    float variance_threshold;
    int iteration;
    DisplayXf xf;
};

RT_DEV float4 f4_mul(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
RT_DEV float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
RT_DEV float4 f4_max(float4 a, float4 b) { // _mm_max_ps(a, b)
    return make_float4(sse_max(a.x, b.x), sse_max(a.y, b.y), sse_max(a.z, b.z), sse_max(a.w, b.w));
}
RT_DEV int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__constant__ float kGaussWeights[5] = {0.2270270270f, 0.1945945946f, 0.1216216216f, 0.0540540541f, 0.0162162162f};

// reversible tonemap of the accumulated image + horizontal pass of the variance blur, over the grown region
__global__ void __launch_bounds__(256) k_nlm_prep(NlmParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.ew * p.eh) {
        return;
    }
    const int x = idx % p.ew, y = idx / p.ew;
    const int xx = p.ex + x, yy = p.ey + y;
    const int w = p.fb.w, h = p.fb.h;
    const int cy = clampi(yy, 0, h - 1);
    const float4 c = p.fb.full[cy * w + clampi(xx, 0, w - 1)];
    const float d = fmaxf(c.x, fmaxf(c.y, c.z)) + 1.0f;
    p.temp_final[idx] = make_float4(c.x / d, c.y / d, c.z / d, c.w / d);
    const float4 center = p.fb.temp[cy * w + clampi(xx, 0, w - 1)];
    float4 res = f4_mul(center, kGaussWeights[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // (the reference's tap positions, RendererCPU.h:700-703: xx - i + 1 and xx + i + 1)
        res = f4_add(res, f4_mul(p.fb.temp[cy * w + clampi(xx - i + 1, 0, w - 1)], kGaussWeights[i + 1]));
        res = f4_add(res, f4_mul(p.fb.temp[cy * w + clampi(xx + i + 1, 0, w - 1)], kGaussWeights[i + 1]));
    }
    p.var_h[idx] = f4_max(res, center);
}

// vertical pass + required-samples update
__global__ void __launch_bounds__(256) k_nlm_vblur(NlmParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int iw = p.ew - 8, ih = p.eh - 8;
    if (idx >= iw * ih) {
        return;
    }
    const int x = 4 + idx % iw, y = 4 + idx / iw;
    const float4 center = p.var_h[y * p.ew + x];
    float4 res = f4_mul(center, kGaussWeights[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        res = f4_add(res, f4_mul(p.var_h[(y - i + 1) * p.ew + x], kGaussWeights[i + 1]));
        res = f4_add(res, f4_mul(p.var_h[(y + i + 1) * p.ew + x], kGaussWeights[i + 1]));
    }
    res = f4_max(res, center);
    p.var_f[y * p.ew + x] = res;
    const int px = x - kNlmExt, py = y - kNlmExt;
    if (px >= 0 && py >= 0 && px < p.rw && py < p.rh) {
        if ((res.x >= p.variance_threshold) | (res.y >= p.variance_threshold) | (res.z >= p.variance_threshold) |
            (res.w >= p.variance_threshold)) {
            p.fb.required_samples[(p.ry + py) * p.fb.w + (p.rx + px)] = uint16_t(p.iteration + 1);
        }
    }
}

constexpr int kNlmBx = 32, kNlmBy = 8, kNlmHalo = kNlmWindow + kNlmPatch; // 4
constexpr int kNlmTw = kNlmBx + 2 * kNlmHalo, kNlmTh = kNlmBy + 2 * kNlmHalo;

__global__ void __launch_bounds__(kNlmBx *kNlmBy) k_nlm_filter(NlmParams p) {
    __shared__ float4 s_col[kNlmTh][kNlmTw];
    __shared__ float4 s_var[kNlmTh][kNlmTw];
    const int bx0 = blockIdx.x * kNlmBx, by0 = blockIdx.y * kNlmBy; // region-relative origin of the block
    // stage the halo tile: ext coords = region-relative + kNlmExt
    for (int i = threadIdx.y * kNlmBx + threadIdx.x; i < kNlmTw * kNlmTh; i += kNlmBx * kNlmBy) {
        const int tx = i % kNlmTw, ty = i / kNlmTw;
        const int gx = min(bx0 + tx - kNlmHalo + kNlmExt, p.ew - 5), gy = min(by0 + ty - kNlmHalo + kNlmExt, p.eh - 5);
        s_col[ty][tx] = p.temp_final[gy * p.ew + gx];
        s_var[ty][tx] = p.var_f[gy * p.ew + gx];
    }
    __syncthreads();
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    if (x >= p.rw || y >= p.rh) {
        return;
    }
    const int lx = threadIdx.x + kNlmHalo, ly = threadIdx.y + kNlmHalo;
    const int w = p.fb.w, h = p.fb.h;
    const int gx = p.rx + x, gy = p.ry + y;
    const float alpha = 1.0f, damping = 0.45f, f0w = 64.0f, f1w = 32.0f;
    const float4 if0 = p.fb.base_color[clampi(gy, 0, h - 1) * w + clampi(gx, 0, w - 1)];
    const float4 if1 = p.fb.depth_normals[clampi(gy, 0, h - 1) * w + clampi(gx, 0, w - 1)];
    float4 sum_output = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float sum_weight = 0.0f;
#pragma unroll 1
    for (int k = -kNlmWindow; k <= kNlmWindow; ++k) {
#pragma unroll 1
        for (int l = -kNlmWindow; l <= kNlmWindow; ++l) {
            float4 cd = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int q = -kNlmPatch; q <= kNlmPatch; ++q) {
#pragma unroll
                for (int pp = -kNlmPatch; pp <= kNlmPatch; ++pp) {
                    const float4 ipx = s_col[ly + q][lx + pp], jpx = s_col[ly + k + q][lx + l + pp];
                    const float4 ivar = s_var[ly + q][lx + pp], jvar = s_var[ly + k + q][lx + l + pp];
#define RT_NLM_LANE(c)                                                                                                 \
    {                                                                                                                  \
        const float mv = sse_min(ivar.c, jvar.c);                                                                      \
        const float df = ipx.c - jpx.c;                                                                                \
        cd.c += (df * df - alpha * (ivar.c + mv)) / (0.0001f + (damping * damping) * (ivar.c + jvar.c));                \
    }
                    RT_NLM_LANE(x)
                    RT_NLM_LANE(y)
                    RT_NLM_LANE(z)
                    RT_NLM_LANE(w)
#undef RT_NLM_LANE
                }
            }
            const float patch_distance = (0.25f * 9.0f) * (cd.x + cd.y + cd.z + cd.w);
            float weight = libm_expf(-fmaxf(0.0f, patch_distance));
            {
                const int jx = clampi(gx + l, 0, w - 1), jy = clampi(gy + k, 0, h - 1);
                const float4 jf0 = p.fb.base_color[jy * w + jx], jf1 = p.fb.depth_normals[jy * w + jx];
                float4 fd;
#define RT_NLM_FEAT(c)                                                                                                 \
    {                                                                                                                  \
        const float d0 = if0.c - jf0.c, d1 = if1.c - jf1.c;                                                            \
        fd.c = sse_max((f0w * d0) * d0, (f1w * d1) * d1);                                                              \
    }
                RT_NLM_FEAT(x)
                RT_NLM_FEAT(y)
                RT_NLM_FEAT(z)
                RT_NLM_FEAT(w)
#undef RT_NLM_FEAT
                const float feature_patch_distance = 0.25f * (fd.x + fd.y + fd.z + fd.w);
                const float feature_weight = libm_expf(-fmaxf(0.0f, fminf(10000.0f, feature_patch_distance)));
                weight = fminf(weight, feature_weight);
            }
            const float4 jc = s_col[ly + k][lx + l];
            sum_output.x += jc.x * weight;
            sum_output.y += jc.y * weight;
            sum_output.z += jc.z * weight;
            sum_output.w += jc.w * weight;
            sum_weight += weight;
        }
    }
    if (sum_weight != 0.0f) {
        sum_output.x /= sum_weight;
        sum_output.y /= sum_weight;
        sum_output.z /= sum_weight;
        sum_output.w /= sum_weight;
    }
    // reversible_tonemap_invert, then Tonemap (Standard) -- RendererCPU.h:771-779
    const float di = 1.0f - fmaxf(sum_output.x, fmaxf(sum_output.y, sum_output.z));
    const float4 col = make_float4(sum_output.x / di, sum_output.y / di, sum_output.z / di, sum_output.w / di);
    const int pix = gy * w + gx;
    p.fb.raw[pix] = col;
    float4 c = col;
    display_transform(p.xf, c);
    c.x = sse_max(0.0f, sse_min(c.x, 1.0f));
    c.y = sse_max(0.0f, sse_min(c.y, 1.0f));
    c.z = sse_max(0.0f, sse_min(c.z, 1.0f));
    c.w = sse_max(0.0f, sse_min(c.w, 1.0f));
    p.fb.final[pix] = c;
}

} // namespace rt
