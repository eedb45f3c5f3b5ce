// rt_unet.cuh -- RendererBase::DenoiseImage(int pass, const RegionContext &): the OIDN-weights UNet denoiser
// (second half).
//
// Behavioural spec: reference internal/RendererCPU.h:790-1007 (the 16 passes), internal/Convolution.h (3x3 convolution
// with zero padding, bias, ReLU on every layer; pre-ops HDRTransfer / PositiveNormalize on the input features, 2x2 max
// pooling after the encoder convolutions, nearest 2x up-sampling + concatenation of the skip tensor in the decoder,
// inverse HDR transfer after the last layer), internal/UNetFilter.cpp (network shape).  Weights are the caller's: 16
// layers of OIHW fp16 weights + fp16 biases (the reference keeps OIDN's `hdr_alb_nrm` set in
// internal/precomputed/__oidn_weights_hdr_alb_nrm.inl; they cross the C-ABI through rc_unet_set_weights).
//
//   pass  layer        input (channels @ scale)                          -> output
//   0     enc_conv0    {HDR(colour) 3, albedo 3, 0.5 n + 0.5 3} @1        -> 32 @1
//   1     enc_conv1    32 @1                                  + pool      -> 32 @1/2
//   2     enc_conv2    32 @1/2                                + pool      -> 48 @1/4
//   3     enc_conv3    48 @1/4                                + pool      -> 64 @1/8
//   4     enc_conv4    64 @1/8                                + pool      -> 80 @1/16
//   5,6   enc_conv5a/b 80 -> 96 -> 96 @1/16
//   7,8   dec_conv4a/b up(96) ++ 64 @1/8  -> 112 -> 112
//   9,10  dec_conv3a/b up(112) ++ 48 @1/4 -> 96 -> 96
//   11,12 dec_conv2a/b up(96) ++ 32 @1/2  -> 64 -> 64
//   13,14 dec_conv1a/b up(64) ++ input 9 @1 -> 64 -> 32
//   15    dec_conv0    32 @1 -> 3, ReLU, inverse HDR transfer -> RAW plane, display transform -> FINAL plane
//
// The network runs on the frame rounded up to a multiple of 16 in both directions (features are zero outside the
// frame, activations are computed there like the reference does); tensors are NHWC without a border, the padding is
// a bounds check.  Two arithmetic paths compute the same layers:
//   * k_unet_conv_f32   fp32 activations and weights (the fp16 weights converted exactly), FFMA: the parity anchor
//   * rt_unet_tc.cuh    fp16 activations, fp32 accumulation on the 5th-generation tensor cores (tcgen05 + TMEM + TMA)
#pragma once

#include <cuda_fp16.h>

#include "rt_kernels.cuh"

namespace rt {

constexpr int kUNetLayers = 16;
constexpr int kUNetInCh = 9;

struct UNetLayerShape {
    int cin1, cin2, cout; // cin1: channels of the (possibly up-sampled) main input, cin2: of the concatenated skip tensor
    int level;            // log2 of the down-scale of the OUTPUT-side convolution grid (before pooling)
    bool up, pool;
};

// the 16 convolutions in pass order
__host__ __device__ inline UNetLayerShape unet_layer(int i) {
    const UNetLayerShape L[kUNetLayers] = {
        {9, 0, 32, 0, false, false},   {32, 0, 32, 0, false, true},  {32, 0, 48, 1, false, true},
        {48, 0, 64, 2, false, true},   {64, 0, 80, 3, false, true},  {80, 0, 96, 4, false, false},
        {96, 0, 96, 4, false, false},  {96, 64, 112, 3, true, false}, {112, 0, 112, 3, false, false},
        {112, 48, 96, 2, true, false}, {96, 0, 96, 2, false, false},  {96, 32, 64, 1, true, false},
        {64, 0, 64, 1, false, false},  {64, 9, 64, 0, true, false},   {64, 0, 32, 0, false, false},
        {32, 0, 3, 0, false, false}};
    return L[i];
}

// HDR transfer function of the network's colour input / output (Convolution.h:62-118)
namespace unet_tf {
constexpr float a = 1.41283765e+03f, b = 1.64593172e+00f, c = 4.31384981e-01f, d = -2.94139609e-03f, e = 1.92653254e-01f,
                f = 6.26026094e-03f, g = 9.98620152e-01f, y0 = 1.57945760e-06f, y1 = 3.22087631e-02f, x0 = 2.23151711e-03f,
                x1 = 3.70974749e-01f;
RT_DEV float input_hdr(float val) {
    const float norm_scale = 0.318967164f;
    if (val <= y0) {
        return a * val * norm_scale;
    } else if (val <= y1) {
        return (b * libm_powf(val, c) + d) * norm_scale;
    }
    return (e * libm_logf(val + f) + g) * norm_scale;
}
RT_DEV float output_hdr(float val) {
    const float norm_scale = 3.13511896f;
    val *= norm_scale;
    if (val <= x0) {
        return val / a;
    } else if (val <= x1) {
        return libm_powf((val - d) / b, 1.0f / c);
    }
    return libm_expf((val - g) / e) - f;
}
} // namespace unet_tf

// the 9 input features of pixel (x, y) of the rounded frame: zero outside the real frame (Convolution.h:226-247)
RT_DEV void unet_features(const FrameBufs &fb, int x, int y, float out[kUNetInCh]) {
    if (x < 0 || y < 0 || x >= fb.w || y >= fb.h) {
#pragma unroll
        for (int i = 0; i < kUNetInCh; ++i) {
            out[i] = 0.0f;
        }
        return;
    }
    const float4 c = fb.full[y * fb.w + x], al = fb.base_color[y * fb.w + x], dn = fb.depth_normals[y * fb.w + x];
    out[0] = unet_tf::input_hdr(c.x);
    out[1] = unet_tf::input_hdr(c.y);
    out[2] = unet_tf::input_hdr(c.z);
    out[3] = al.x;
    out[4] = al.y;
    out[5] = al.z;
    out[6] = 0.5f * dn.x + 0.5f;
    out[7] = 0.5f * dn.y + 0.5f;
    out[8] = 0.5f * dn.z + 0.5f;
}

struct UNetConvParams {
    const float *in1; // NHWC, cin1 channels, grid (w1 x h1) = conv grid, or half of it when `up`
    const float *in2; // NHWC, cin2 channels on the conv grid (skip tensor), or null; for pass 13 the features come from fb
    const float *weights; // [cout][9][cin1 + cin2]  (tap-major, input channels contiguous)
    const float *bias;    // [cout]
    float *out;           // NHWC, cout channels, conv grid or half of it when `pool`
    FrameBufs fb;         // passes 0, 13, 15
    int cin1, cin2, cout;
    int w, h;             // conv grid (rounded frame >> level)
    int rx, ry, rw, rh;   // region of the conv grid to compute
    int up, pool, feat_in2, feat_in1, last;
    DisplayXf xf;
};

constexpr int kConvTile = 8;      // 8 x 8 output pixels per block
constexpr int kConvCoutBlk = 16;  // output channels per block
constexpr int kConvCinBlk = 16;   // input channels staged per step

// Direct 3x3 convolution, fp32.  Block = 64 threads = an 8 x 8 pixel tile x 16 output channels; input channels are staged
// 16 at a time through shared memory together with their weights (a 10 x 10 halo tile and a [16][9][16] weight slab).
__global__ void __launch_bounds__(64) k_unet_conv_f32(UNetConvParams p) {
    __shared__ float s_in[kConvTile + 2][kConvTile + 2][kConvCinBlk + 1];
    __shared__ float s_w[kConvCoutBlk][9][kConvCinBlk];
    const int tx = threadIdx.x % kConvTile, ty = threadIdx.x / kConvTile;
    const int x0 = p.rx + blockIdx.x * kConvTile, y0 = p.ry + blockIdx.y * kConvTile;
    const int co0 = blockIdx.z * kConvCoutBlk;
    const int cin = p.cin1 + p.cin2;
    float acc[kConvCoutBlk];
#pragma unroll
    for (int i = 0; i < kConvCoutBlk; ++i) {
        acc[i] = (co0 + i < p.cout) ? p.bias[co0 + i] : 0.0f;
    }
    for (int c0 = 0; c0 < cin; c0 += kConvCinBlk) {
        __syncthreads();
        // stage the halo tile of input channels [c0, c0 + 16)
        for (int i = threadIdx.x; i < (kConvTile + 2) * (kConvTile + 2); i += 64) {
            const int hx = i % (kConvTile + 2), hy = i / (kConvTile + 2);
            const int x = x0 + hx - 1, y = y0 + hy - 1;
            const bool inside = x >= 0 && y >= 0 && x < p.w && y < p.h;
            float feat[kUNetInCh];
            const bool need_feat = (p.feat_in1 && c0 < p.cin1) || (p.feat_in2 && c0 + kConvCinBlk > p.cin1);
            if (need_feat) {
                unet_features(p.fb, x, y, feat); // zero outside the real frame
            }
            for (int k = 0; k < kConvCinBlk; ++k) {
                const int c = c0 + k;
                float v = 0.0f;
                if (inside && c < cin) {
                    if (c < p.cin1) {
                        if (p.feat_in1) {
                            v = feat[c];
                        } else if (p.up) {
                            v = p.in1[(size_t(y >> 1) * (p.w >> 1) + (x >> 1)) * p.cin1 + c];
                        } else {
                            v = p.in1[(size_t(y) * p.w + x) * p.cin1 + c];
                        }
                    } else if (p.feat_in2) {
                        v = feat[c - p.cin1];
                    } else {
                        v = p.in2[(size_t(y) * p.w + x) * p.cin2 + (c - p.cin1)];
                    }
                }
                s_in[hy][hx][k] = v;
            }
        }
        for (int i = threadIdx.x; i < kConvCoutBlk * 9 * kConvCinBlk; i += 64) {
            const int k = i % kConvCinBlk, t = (i / kConvCinBlk) % 9, o = i / (kConvCinBlk * 9);
            const int c = c0 + k, co = co0 + o;
            s_w[o][t][k] = (c < cin && co < p.cout) ? p.weights[(size_t(co) * 9 + t) * cin + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t % 3;
#pragma unroll 4
            for (int k = 0; k < kConvCinBlk; ++k) {
                const float v = s_in[ty + dy][tx + dx][k];
#pragma unroll
                for (int o = 0; o < kConvCoutBlk; ++o) {
                    acc[o] = __fmaf_rn(v, s_w[o][t][k], acc[o]);
                }
            }
        }
    }
    const int x = x0 + tx, y = y0 + ty;
    if (x >= p.rx + p.rw || y >= p.ry + p.rh) {
        // (threads outside the region still took part in the staging above)
        if (!p.pool) {
            return;
        }
    }
#pragma unroll
    for (int o = 0; o < kConvCoutBlk; ++o) {
        acc[o] = fmaxf(0.0f, acc[o]); // ReLU on every layer
    }
    if (p.pool) {
        // 2 x 2 max over the tile's pixel quads: the quad lives in lanes {l, l + 1, l + 8, l + 9}
#pragma unroll
        for (int o = 0; o < kConvCoutBlk; ++o) {
            float m = fmaxf(acc[o], __shfl_down_sync(0xffffffffu, acc[o], 1));
            m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, kConvTile));
            acc[o] = m;
        }
        if ((tx & 1) || (ty & 1) || x >= p.rx + p.rw || y >= p.ry + p.rh) {
            return;
        }
        float *dst = p.out + (size_t(y >> 1) * (p.w >> 1) + (x >> 1)) * p.cout + co0;
#pragma unroll
        for (int o = 0; o < kConvCoutBlk; ++o) {
            if (co0 + o < p.cout) {
                dst[o] = acc[o];
            }
        }
        return;
    }
    if (p.last) {
        // dec_conv0: inverse HDR transfer into RAW, display transform into FINAL (RendererCPU.h:975-996)
        if (x < p.fb.w && y < p.fb.h && co0 == 0) {
            const int pix = y * p.fb.w + x;
            const float4 full = p.fb.full[pix];
            float4 c = make_float4(unet_tf::output_hdr(acc[0]), unet_tf::output_hdr(acc[1]), unet_tf::output_hdr(acc[2]), full.w);
            p.fb.raw[pix] = c;
            display_transform(p.xf, c);
            c.x = sse_max(0.0f, sse_min(c.x, 1.0f));
            c.y = sse_max(0.0f, sse_min(c.y, 1.0f));
            c.z = sse_max(0.0f, sse_min(c.z, 1.0f));
            c.w = sse_max(0.0f, sse_min(c.w, 1.0f));
            p.fb.final[pix] = c;
        }
        return;
    }
    float *dst = p.out + (size_t(y) * p.w + x) * p.cout + co0;
#pragma unroll
    for (int o = 0; o < kConvCoutBlk; ++o) {
        if (co0 + o < p.cout) {
            dst[o] = acc[o];
        }
    }
}

} // namespace rt
