// rt_unet_tc.cuh -- the UNet's 3x3 convolutions as implicit GEMMs on the 5th-generation tensor cores (sm_100a):
// tcgen05.mma with the accumulator in TMEM, operands staged by TMA (cp.async.bulk.tensor) into 128-byte-swizzled shared
// memory, an mbarrier pipeline between one TMA-producer thread, one MMA-issuer thread and four epilogue warps.
// This is the one dense contraction on the path; the reference's counterpart is the
// cooperative-matrix shader internal/shaders/convolution.comp.glsl.  fp16 operands (the network's weights ARE fp16;
// activations are rounded to fp16 between layers like the reference's GPU path), fp32 accumulation.
//
// Data layout (device only): every activation tensor is NHWC fp16 with a one-pixel zero border and a channel stride
// padded to a multiple of 64: element (y, x, c) of an H x W grid sits at ((y + 1) (W + 2) + (x + 1)) Cs + c.  Padded
// channels and the border are zero and are never written, so a convolution tap is a plain window of this array: the
// A operand of tap (dy, dx) for the 128 output pixels (y, x0 .. x0 + 127) is the 2-D box [128 pixels][64 channels]
// starting at pixel (y + dy) (W + 2) + x0 + dx -- one TMA load, no im2col, zero padding for free.  Weights are stored
// [tap][cout padded to 16][Cs] fp16: the B operand of (tap, 64-channel block) is the box [cout][64].
//
// GEMM per tile: D[128 pixels][N = cout] += sum over 9 taps x K blocks of A[128][64] B[N][64]^T, K = 16 per tcgen05.mma,
// M = 128, D in N TMEM columns (fp32).  The K range may span TWO input tensors: a decoder convolution reads the
// up-sampled previous tensor in its first blocks and the skip tensor in the following ones (the concatenation is never
// materialised), and the layer BEFORE a decoder convolution stores each output pixel to the 2 x 2 block of the finer
// grid (nearest up-sampling fused into its epilogue).  Epilogue: tcgen05.ld -> bias + ReLU -> fp16 -> NHWC store (or,
// for the last layer, inverse HDR transfer + display transform into the RAW / FINAL planes).
// CTAs are persistent (2 per SM), each with two TMEM accumulators: the epilogue of one tile overlaps the TMA loads and
// MMAs of the next.  Per image row of the 3 x 3 window ONE haloed activation box (130 pixels) is loaded and multiplied
// three times, shifted by one pixel per horizontal tap (the 128-byte swizzle is a function of the absolute shared
// address, so a shifted start address is all the descriptor needs), and a layer whose whole filter fits next to the
// activation ring keeps it resident for the CTA's lifetime.  1080p, all 16 layers: 2.1 ms (first build, one CTA per tile
// and separate gather passes: 4.7 ms; fp32 path: 31 ms).  Measured and not kept: weights resident with ONE CTA per SM
// for the 147 KB filter of dec_conv1a (fewer bytes in flight per SM: slower), L2 tensor prefetch one tile ahead (no gain:
// the loads are bound by L2 -> SMEM throughput, ~11 TB/s, not by DRAM latency), 64 / 32-byte swizzled boxes for the thin
// tensors (correct, not faster).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include "rt_unet.cuh"

namespace rt {
namespace tc {

constexpr int kMaxStages = 8;            // B ring slots (one tap of one K block each)
constexpr int kMaxASlots = 6;            // A ring slots (one haloed row segment of one K block each)
constexpr int kHaloRows = 130;           // 128 output pixels + one neighbour each side: serves the three horizontal taps
constexpr int kASlotBytes = 17 * 1024;   // 130 rows x 128 B, rounded up to the 1024-byte swizzle period
constexpr int kTileM = 128;              // output pixels per tile (one row segment)
constexpr int kBlockK = 64;              // channels per pipeline stage (one 128-byte swizzle row)
constexpr int kThreads = 192;            // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kATileBytes = kTileM * kBlockK * 2;
constexpr int kSmemBudget = 110 * 1024;  // per CTA: two CTAs per SM (their 2 x 2 accumulators fill the 512 TMEM columns)

RT_DEV uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

RT_DEV void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
RT_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
RT_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
RT_DEV void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
RT_DEV void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
RT_DEV void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
RT_DEV void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
RT_DEV void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "setp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
                 "}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// K-major operand tile, rows of 128 bytes, 128-byte swizzle, 8-row groups 1024 bytes apart (cute/arch/mma_sm100_desc.hpp)
// `row_shift` (0..7): the tile starts that many 128-byte rows past a 1024-byte boundary of the swizzle pattern (the haloed
// A tile is read three times, shifted by one pixel row per horizontal tap).  Measured on B200: the unit applies the
// 128-byte swizzle to the ABSOLUTE shared-memory address, so moving the start address is all it takes -- with the
// descriptor's base-offset field set to the shift the results are wrong, with 0 they are right (use_base_offset stays as
// the switch that experiment used)
RT_DEV uint64_t umma_desc_sw128(const void *smem_tile, uint32_t row_shift = 0, uint32_t use_base_offset = 0) {
    const uint64_t addr = uint64_t(smem_u32(smem_tile)) + row_shift * 128u;
    return ((addr >> 4) & 0x3fffull) | (1ull << 16) /* LBO (ignored for swizzled K-major) */ |
           (uint64_t(1024 >> 4) << 32) /* SBO */ | (1ull << 46) /* descriptor version of sm_100 */ |
           (uint64_t(use_base_offset ? (row_shift & 7u) : 0u) << 49) /* matrix base offset */ | (2ull << 61) /* SWIZZLE_128B */;
}
// kind::f16 instruction descriptor: D = fp32, A = B = fp16, both K-major, M = 128, N
RT_DEV uint32_t umma_idesc_f16(int n) { return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(kTileM >> 4) << 24); }

struct ConvTcParams {
    const float *bias;   // [n_pad] fp32 (zero for padded channels)
    __half *out;         // bordered NHWC fp16, channel stride out_cs; null for the last layer
    FrameBufs fb;        // last layer: RAW / FINAL planes
    int w, h;            // convolution grid
    int cin1, nkb1;      // first input tensor: real channels (multiple of 16) and 64-channel blocks
    int cin2, nkb2;      // second input tensor (decoder skip connection), 0 blocks when absent
    int n;               // output channels padded to 16 (= UMMA N)
    int cout;            // real output channels
    int out_cs;
    int up;              // write every output pixel to the 2 x 2 block of a (2w x 2h) tensor: nearest up-sampling fused
    int tmem_cols;       // power of two >= max(32, n); the CTA allocates two accumulators
    int stages;          // B ring depth that fits kSmemBudget next to the A ring
    int base_offset;     // descriptor base-offset field carries the row shift (experiment switch, 1 = on)
    int b_resident;      // the layer's whole weight tensor (9 taps x K blocks) stays in shared memory for the CTA's lifetime
    int a_slots;         // A ring depth
    int tiles_x, tiles;  // row segments per image row, total
    int last;
    int rx, ry, rw, rh;  // last layer: frame rect whose pixels are written
    DisplayXf xf;
};

// Persistent CTAs: tile t = blockIdx.x + k gridDim.x.  The TMA producer runs ahead across tile boundaries; the MMA issuer
// alternates between two TMEM accumulators, so the epilogue of tile k (TMEM -> bias/ReLU -> fp16 -> global) overlaps the
// loads and MMAs of tile k + 1.  dynamic shared memory: stages x (A tile 16 KB + B tile n x 128 B), 1024-byte aligned.
__global__ void __launch_bounds__(kThreads) k_unet_conv_tc(const __grid_constant__ CUtensorMap map_a1,
                                                           const __grid_constant__ CUtensorMap map_a2,
                                                           const __grid_constant__ CUtensorMap map_b, ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the 128-byte swizzle pattern repeats every 1024 bytes of SHARED address: align the stage buffers in that space
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t a_full[kMaxASlots], a_empty[kMaxASlots], b_full[kMaxStages], b_empty[kMaxStages], b_all, acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_tile_bytes = p.n * kBlockK * 2;
    const int b_slot_bytes = (b_tile_bytes + 1023) & ~1023;
    const int AS = p.a_slots;
    uint8_t *smem_a = smem, *smem_b = smem + AS * kASlotBytes;
    const int nkb = p.nkb1 + p.nkb2;
    const int S = p.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < AS; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        mbar_init(&b_all, 1);
        for (int s = 0; s < S; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4); // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"(uint32_t(2 * p.tmem_cols)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    // K loop of one tile: for every image row dy of the 3 x 3 window and every 64-channel block ONE haloed A box
    // (130 pixels) is loaded and multiplied three times, shifted by one pixel per horizontal tap, against the three
    // B boxes of that row of the filter -- the activations cross L2 -> SMEM 3 times per layer instead of 9
    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer ----
            uint32_t ia = 0, ib = 0;
            if (p.b_resident) {
                // the whole filter (9 taps x K blocks) once; the tile loop then streams activations only
                mbar_expect_tx(&b_all, uint32_t(9 * nkb * b_tile_bytes));
                for (int tap = 0; tap < 9; ++tap) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        tma_load_2d(smem_b + (tap * nkb + kb) * b_slot_bytes, &map_b, &b_all, kb * kBlockK, tap * p.n);
                    }
                }
            }
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                const int y = tile / p.tiles_x, x0 = (tile % p.tiles_x) * kTileM;
                for (int dy = 0; dy < 3; ++dy) {
                    for (int kb = 0; kb < nkb; ++kb, ++ia) {
                        const uint32_t aslot = ia % uint32_t(AS), around = ia / uint32_t(AS);
                        if (around > 0) {
                            mbar_wait(&a_empty[aslot], (around - 1) & 1);
                        }
                        mbar_expect_tx(&a_full[aslot], uint32_t(kHaloRows * kBlockK * 2));
                        const int pix = (y + dy) * (p.w + 2) + x0;
                        if (kb < p.nkb1) {
                            tma_load_2d(smem_a + aslot * kASlotBytes, &map_a1, &a_full[aslot], kb * kBlockK, pix);
                        } else {
                            tma_load_2d(smem_a + aslot * kASlotBytes, &map_a2, &a_full[aslot], (kb - p.nkb1) * kBlockK, pix);
                        }
                        for (int dx = 0; dx < 3 && !p.b_resident; ++dx, ++ib) {
                            const uint32_t bslot = ib % uint32_t(S), bround = ib / uint32_t(S);
                            if (bround > 0) {
                                mbar_wait(&b_empty[bslot], (bround - 1) & 1);
                            }
                            mbar_expect_tx(&b_full[bslot], uint32_t(b_tile_bytes));
                            tma_load_2d(smem_b + bslot * b_slot_bytes, &map_b, &b_full[bslot], kb * kBlockK, (dy * 3 + dx) * p.n);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer ----
            const uint32_t idesc = umma_idesc_f16(p.n);
            uint32_t ia = 0, ib = 0, k = 0;
            if (p.b_resident) {
                mbar_wait(&b_all, 0);
            }
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++k) {
                const uint32_t buf = k & 1u;
                if (k >= 2) {
                    mbar_wait(&acc_empty[buf], ((k >> 1) - 1) & 1); // the epilogue has drained this accumulator
                    tcgen05_fence_after();
                }
                const uint32_t tmem_d = tmem_base + buf * uint32_t(p.tmem_cols);
                uint32_t accumulate = 0;
                for (int dy = 0; dy < 3; ++dy) {
                    for (int kb = 0; kb < nkb; ++kb, ++ia) {
                        const uint32_t aslot = ia % uint32_t(AS), around = ia / uint32_t(AS);
                        mbar_wait(&a_full[aslot], around & 1);
                        tcgen05_fence_after();
                        const int kcount =
                            (kb < p.nkb1 ? min(kBlockK, p.cin1 - kb * kBlockK) : min(kBlockK, p.cin2 - (kb - p.nkb1) * kBlockK)) / 16;
                        const uint8_t *a = smem_a + aslot * kASlotBytes;
                        for (int dx = 0; dx < 3; ++dx) {
                            uint32_t bslot;
                            if (p.b_resident) {
                                bslot = uint32_t((dy * 3 + dx) * nkb + kb);
                            } else {
                                bslot = ib % uint32_t(S);
                                mbar_wait(&b_full[bslot], (ib / uint32_t(S)) & 1);
                                tcgen05_fence_after();
                                ++ib;
                            }
                            const uint64_t adesc = umma_desc_sw128(a, uint32_t(dx), uint32_t(p.base_offset));
                            const uint64_t bdesc = umma_desc_sw128(smem_b + bslot * b_slot_bytes);
                            for (int kk = 0; kk < kcount; ++kk) {
                                // 16 fp16 = 32 bytes further along the swizzled 128-byte row: start-address field += 2
                                umma_f16(tmem_d, adesc + uint64_t(2 * kk), bdesc + uint64_t(2 * kk), idesc, accumulate);
                                accumulate = 1;
                            }
                            if (!p.b_resident) {
                                umma_commit(&b_empty[bslot]); // frees the B slot once these MMAs have read it
                            }
                        }
                        umma_commit(&a_empty[aslot]); // all three taps have read the haloed A box
                    }
                }
                umma_commit(&acc_full[buf]); // accumulator complete
            }
        }
    } else {
        // ---- epilogue: warps 2..5; a warp may only touch the TMEM lanes [32 (warp % 4), +32) ----
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane; // GEMM row = pixel x0 + row
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++k) {
            const int y = tile / p.tiles_x, x0 = (tile % p.tiles_x) * kTileM;
            const int x = x0 + row;
            const uint32_t buf = k & 1u;
            mbar_wait(&acc_full[buf], (k >> 1) & 1);
            tcgen05_fence_after();
            float outv[3] = {0.0f, 0.0f, 0.0f};
            for (int c0 = 0; c0 < p.n; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + buf * uint32_t(p.tmem_cols) + (uint32_t(lane_base) << 16) + uint32_t(c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
                             "[%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p.last) {
                    if (c0 == 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            outv[i] = fmaxf(0.0f, __uint_as_float(r[i]) + p.bias[i]);
                        }
                    }
                    continue;
                }
                if (x < p.w) {
                    __half2 h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float v0 = fmaxf(0.0f, __uint_as_float(r[2 * i]) + p.bias[c0 + 2 * i]);
                        const float v1 = fmaxf(0.0f, __uint_as_float(r[2 * i + 1]) + p.bias[c0 + 2 * i + 1]);
                        h[i] = __floats2half2_rn(v0, v1);
                    }
                    // padded output channels (>= cout) have zero weights and zero bias: they are written as ReLU(0) = 0
                    const uint4 lo = *reinterpret_cast<const uint4 *>(&h[0]), hi = *reinterpret_cast<const uint4 *>(&h[4]);
                    if (p.up) {
                        const size_t pitch = size_t(2 * p.w + 2);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 *dst = reinterpret_cast<uint4 *>(
                                p.out + ((size_t(2 * y + 1 + (q >> 1)) * pitch + size_t(2 * x + 1 + (q & 1))) * p.out_cs + c0));
                            dst[0] = lo;
                            dst[1] = hi;
                        }
                    } else {
                        uint4 *dst = reinterpret_cast<uint4 *>(p.out + (size_t(y + 1) * (p.w + 2) + size_t(x + 1)) * p.out_cs + c0);
                        dst[0] = lo;
                        dst[1] = hi;
                    }
                }
            }
            // this warp has read its lanes of the accumulator: hand it back to the MMA issuer
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
            }
            if (p.last && x >= p.rx && x < p.rx + p.rw && y >= p.ry && y < p.ry + p.rh) {
                const int pix = y * p.fb.w + x;
                const float4 full = p.fb.full[pix];
                float4 c = make_float4(unet_tf::output_hdr(outv[0]), unet_tf::output_hdr(outv[1]), unet_tf::output_hdr(outv[2]), full.w);
                p.fb.raw[pix] = c;
                display_transform(p.xf, c);
                c.x = sse_max(0.0f, sse_min(c.x, 1.0f));
                c.y = sse_max(0.0f, sse_min(c.y, 1.0f));
                c.z = sse_max(0.0f, sse_min(c.z, 1.0f));
                c.w = sse_max(0.0f, sse_min(c.w, 1.0f));
                p.fb.final[pix] = c;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(2 * p.tmem_cols)));
    }
}

// ---- the element-wise helpers around the convolutions (fp16 bordered tensors) -----------------------------------------
// network input: 9 features in a 64-channel-stride tensor (channels 9..63 stay zero)
__global__ void k_unet_feat_h(FrameBufs fb, __half *out, int w, int h, int cs) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) {
        return;
    }
    float f[kUNetInCh];
    unet_features(fb, x, y, f);
    __half2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = __floats2half2_rn(2 * i < kUNetInCh ? f[2 * i] : 0.0f, 2 * i + 1 < kUNetInCh ? f[2 * i + 1] : 0.0f);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t(y + 1) * (w + 2) + size_t(x + 1)) * cs);
    dst[0] = *reinterpret_cast<const uint4 *>(&v[0]);
    dst[1] = *reinterpret_cast<const uint4 *>(&v[4]);
}

// 2 x 2 max pooling of a full-resolution tensor into the half-resolution one: one thread per output pixel and 8 channels
__global__ void k_unet_pool_h(const __half *__restrict__ in, int in_cs, __half *__restrict__ out, int out_cs, int c8, int w, int h) {
    const int wo = w >> 1, ho = h >> 1;
    const size_t gid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int g = int(gid % size_t(c8));
    const size_t pix = gid / size_t(c8);
    if (pix >= size_t(wo) * ho) {
        return;
    }
    const int x = int(pix % size_t(wo)), y = int(pix / size_t(wo));
    const __half *p0 = in + (size_t(2 * y + 1) * (w + 2) + size_t(2 * x + 1)) * in_cs + g * 8;
    const __half *p1 = p0 + size_t(w + 2) * in_cs;
    const uint4 a = *reinterpret_cast<const uint4 *>(p0), b = *reinterpret_cast<const uint4 *>(p0 + in_cs),
                c = *reinterpret_cast<const uint4 *>(p1), d = *reinterpret_cast<const uint4 *>(p1 + in_cs);
    const __half2 *ha = reinterpret_cast<const __half2 *>(&a), *hb = reinterpret_cast<const __half2 *>(&b),
                  *hc = reinterpret_cast<const __half2 *>(&c), *hd = reinterpret_cast<const __half2 *>(&d);
    uint4 r;
    __half2 *hr = reinterpret_cast<__half2 *>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hr[i] = __hmax2(__hmax2(ha[i], hb[i]), __hmax2(hc[i], hd[i]));
    }
    *reinterpret_cast<uint4 *>(out + (size_t(y + 1) * (wo + 2) + size_t(x + 1)) * out_cs + g * 8) = r;
}

} // namespace tc
} // namespace rt
