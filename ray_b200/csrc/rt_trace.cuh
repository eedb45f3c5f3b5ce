// rt_trace.cuh -- the persistent trace kernels (closest hit + any hit) as a warp-synchronous PHASE machine.
//
// Behavioural spec (unchanged, see rt_traverse.cuh): reference internal/CoreRef.cpp
//   Traverse_TLAS/BLAS_WithStack_ClosestHit (BVH8) :2027-2133, :2495-2578     ..._AnyHit :2282-2396, :2695-2787
//   IntersectScene (closest, transparency loop)    :3041-3158                 IntersectScene(shadow)  :3160-3262
// The visit ORDER of every ray is the reference's (ordered push, sort_top3/4/N, pop nearest, cull on pop), so every
// tie resolves as in the reference.  What changed against the first-round kernels is HOW a warp walks its 32 rays:
//
//   * one lane = one ray, but the loop body is a sequence of phases every lane passes in step:
//         [epilogue + refill] -> [inner node] -> [BLAS leaf] -> [TLAS leaf / instance] -> [pop]
//     A lane executes the phase its ray needs and idles through the others; the code of a phase is therefore issued once
//     for all lanes that need it instead of once per divergent path (round 1: 10.8 / 6.7 of 32 lanes per instruction).
//   * lanes whose ray has finished take the next ray from the list's queue head (warp-aggregated atomic) instead of
//     waiting for the slowest ray of a 32-ray packet.
//   * the 8-wide slab test needs no min/max to order the two planes of an axis: the sign of inv_d picks the near and
//     the far plane BY ADDRESS (inv >= 0: near = bbox_min; the products are monotone in the plane, so min(lo, hi) is
//     lo), two children per instruction with the packed-fp32 pipe of sm_100 (FADD2 / FMUL2), 3-input FMNMX3 for the
//     three axes.  ~110 instead of ~220 instructions per node.  Results are the same floats: sub.rn / mul.rn per element.
//   * the traversal stack lives in shared memory ([entry][thread] interleaved, 8-byte entries, kStackSmem deep; deeper
//     entries overflow to a local array that is almost never touched), the hit distances of a node go through a
//     32-byte shared scratch so the sorted push indexes them without a local-memory array.
//   * nodes are read from a device-built copy (k_build_dnodes) whose child words say "leaf, first, blocks" directly (no
//     dependent load to find out a child is a leaf) and whose empty slots carry a box at +inf that no ray can hit (the
//     reference keeps a zero box at the origin and would index node 0x7fffffff if a ray ever hit it).
#pragma once

#include "rt_kernels.cuh"

#ifndef RT_STACK_SMEM
#define RT_STACK_SMEM 8
#endif
#ifndef RT_TRACE_THREADS
#define RT_TRACE_THREADS 128
#endif

namespace rt {

constexpr int kStackSmem = RT_STACK_SMEM;
constexpr int kTraceThreads = RT_TRACE_THREADS;
constexpr uint32_t kNoNode = kEmptyChild;            // "nothing to visit": the lane needs a pop
constexpr uint32_t kLeafFirstBits = (1u << 27) - 1u; // leaf word: kLeafBit | (blocks - 1) << 27 | first primitive
constexpr int kLeafBlocksShift = 27;

// ---- packed binary32 pairs (sm_100: FADD2 / FMUL2; each element is an IEEE round-to-nearest operation) -------------
typedef unsigned long long f2;
RT_DEV f2 pk2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
RT_DEV void upk2(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
RT_DEV f2 bc2(float v) { return pk2(v, v); }
RT_DEV f2 sub2(f2 a, f2 b) {
    f2 c;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    return c;
}
RT_DEV f2 mul2(f2 a, f2 b) {
    f2 c;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    return c;
}

// ---- device node copy ----------------------------------------------------------------------------------------------
// Same 224-byte layout as wbvh_node_t; written once per scene upload from the array that crossed the C-ABI.
//   child word of an inner child : its node index
//   child word of a leaf child   : kLeafBit | (blocks - 1) << 27 | first   (first = first triangle slot of a BLAS leaf,
//                                  mesh-instance index of a TLAS leaf; blocks = 8-triangle blocks the leaf spans)
//   empty slot                   : kEmptyChild, box = a point at +inf (never hit)
// Node slots inside the array's capacity that no live node points at (SparseStorage keeps stale bytes in freed slots) only
// have to be harmless here; whether every node REACHABLE from the TLAS root can be encoded is checked on the host
// (validate_bvh in ray_cuda.cu).  A child box is stored with its planes ordered (min <= max): the reference's slab test
// takes min / max of the two products, so it is symmetric in the two planes, and "near plane by sign" is that same value.
RT_DEV uint32_t leaf_word(uint32_t c0, uint32_t c1) {
    const uint32_t first = c0 & kPrimIndexBits;
    const uint32_t blocks = ((first & 7u) + c1 + 7u) / 8u;
    if (first >= kLeafFirstBits || blocks == 0 || blocks > 16) {
        return kEmptyChild;
    }
    return kLeafBit | ((blocks - 1u) << kLeafBlocksShift) | first;
}

__global__ void k_build_dnodes(const WNode *__restrict__ src, WNode *__restrict__ dst, uint32_t first, uint32_t count) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = first + (gid >> 3), c = gid & 7u; // nodes [first, count): a TLAS-only refresh starts past the BLASes
    if (n >= count) {
        return;
    }
    const WNode &s = src[n];
    WNode &d = dst[n];
    if (s.child[0] & kLeafBit) { // leaf nodes are never read through the copy; keep them verbatim
        for (int a = 0; a < 3; ++a) {
            d.bbox_min[a][c] = s.bbox_min[a][c];
            d.bbox_max[a][c] = s.bbox_max[a][c];
        }
        d.child[c] = s.child[c];
        return;
    }
    const uint32_t ch = s.child[c];
    uint32_t word = kEmptyChild;
    if (ch != kEmptyChild && ch < count) {
        const uint32_t c0 = src[ch].child[0];
        word = (c0 & kLeafBit) ? leaf_word(c0, src[ch].child[1]) : ch;
    }
    if (word == kEmptyChild) {
        const float inf = __int_as_float(0x7f800000);
        for (int a = 0; a < 3; ++a) {
            d.bbox_min[a][c] = inf;
            d.bbox_max[a][c] = inf;
        }
        d.child[c] = kEmptyChild;
        return;
    }
    for (int a = 0; a < 3; ++a) {
        const float lo = s.bbox_min[a][c], hi = s.bbox_max[a][c];
        d.bbox_min[a][c] = fminf(lo, hi);
        d.bbox_max[a][c] = fmaxf(lo, hi);
    }
    d.child[c] = word;
}

// BLAS root word of every mesh instance (the root itself may be a leaf)
__global__ void k_build_blas_roots(const WNode *__restrict__ src, const MeshInstance *__restrict__ inst, uint32_t count,
                                   uint32_t node_count, uint32_t *__restrict__ roots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) {
        return;
    }
    const uint32_t r = inst[i].node_index;
    if (r >= node_count) {
        roots[i] = kEmptyChild; // freed SparseStorage slot: never referenced by a TLAS leaf
        return;
    }
    const uint32_t c0 = src[r].child[0];
    roots[i] = (c0 & kLeafBit) ? leaf_word(c0, src[r].child[1]) : r;
}

// ---- shared-memory working set of a trace block ---------------------------------------------------------------------
// Only what every node visit touches stays in registers (the ray in the current space, t, the node word, the stack
// size); what is read once or twice per ray sits in shared memory next to the stack, or the kernel spills at 80 registers.
struct TraceSmem {
    uint2 stack[kStackSmem][kTraceThreads]; // {node word, bits(dist)}
    float4 dist[2][kTraceThreads];          // tmin of the 8 children of the node being visited
    uint4 cword[2][kTraceThreads];          // their child words (loaded with the boxes: no dependent load after the test)
    float4 winv[kTraceThreads];             // safe_invert of the world-space direction (restored when a BLAS is left)
    float4 wro[kTraceThreads];              // world-space origin of the traversal in flight | kernel-specific word
    float4 wrd[kTraceThreads];              // world-space direction                        | kernel-specific word
    float4 hit[kTraceThreads];              // closest hit so far: {u, v, bits(prim), bits(obj)}
    float4 aux[kTraceThreads];              // shadow kernel: {throughput rgb, -}
    uint32_t idx[kTraceThreads];            // index of the lane's ray in its list
};

struct TStack {
    uint2 ovf[2 * kMaxStack - kStackSmem];
};

#define RT_ST_PUT(sm, ov, j, idx, dst)                                                                                  \
    do {                                                                                                               \
        const int _j = (j);                                                                                            \
        const uint2 _e = make_uint2((idx), __float_as_uint(dst));                                                      \
        if (_j < kStackSmem) {                                                                                         \
            (sm).stack[_j][threadIdx.x] = _e;                                                                          \
        } else {                                                                                                       \
            (ov).ovf[_j - kStackSmem] = _e;                                                                            \
        }                                                                                                              \
    } while (0)

RT_DEV uint2 st_get(const TraceSmem &sm, const TStack &ov, int j) {
    return (j < kStackSmem) ? sm.stack[j][threadIdx.x] : ov.ovf[j - kStackSmem];
}

// near-plane byte offsets of the three axes inside a node: axis a starts at 32 a (bbox_min) / 96 + 32 a (bbox_max)
struct AxisSel {
    uint32_t nx, ny, nz;
};
RT_DEV AxisSel axis_sel(v3 inv_d) {
    AxisSel s;
    s.nx = (__float_as_uint(inv_d.x) >> 31) * 96u;
    s.ny = 32u + (__float_as_uint(inv_d.y) >> 31) * 96u;
    s.nz = 64u + (__float_as_uint(inv_d.z) >> 31) * 96u;
    return s;
}

// 8 slab tests.  Reference: bbox_test_oct, CoreRef.cpp:323-350:  lo = inv_d * (bmin - o), hi = inv_d * (bmax - o),
// tmin = max over axes of min(lo, hi), tmax = min over axes of max(lo, hi), tmax *= 1.00000024f,
// hit = tmin <= tmax & tmin <= t & tmax > 0.  For bmin <= bmax (checked by k_build_dnodes) and inv_d >= 0 the rounded
// products satisfy lo <= hi (subtraction and multiplication by a non-negative number are monotone under rounding),
// for inv_d < 0 hi <= lo, so min(lo, hi) is the product with the plane the sign of inv_d selects -- the same float.
// max / min over the three axes are order-independent for non-NaN operands (no NaN can form: inv_d is finite and
// non-zero by safe_invert, box planes are finite or +inf, inf * nonzero = inf).
RT_DEV uint32_t box8_near_far(const char *__restrict__ node, AxisSel s, v3 o, v3 inv_d, float t, float4 &d0, float4 &d1) {
    const float4 *nxp = reinterpret_cast<const float4 *>(node + s.nx);
    const float4 *nyp = reinterpret_cast<const float4 *>(node + s.ny);
    const float4 *nzp = reinterpret_cast<const float4 *>(node + s.nz);
    const float4 *fxp = reinterpret_cast<const float4 *>(node + (96u - s.nx));
    const float4 *fyp = reinterpret_cast<const float4 *>(node + (160u - s.ny));
    const float4 *fzp = reinterpret_cast<const float4 *>(node + (224u - s.nz));
    const f2 ox = bc2(o.x), oy = bc2(o.y), oz = bc2(o.z);
    const f2 ix = bc2(inv_d.x), iy = bc2(inv_d.y), iz = bc2(inv_d.z);
    const f2 slack = bc2(1.00000024f);
    uint32_t mask = 0;
    float tmin[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 ax = __ldg(nxp + h), ay = __ldg(nyp + h), az = __ldg(nzp + h);
        const float4 bx = __ldg(fxp + h), by = __ldg(fyp + h), bz = __ldg(fzp + h);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const f2 nx = mul2(ix, sub2(q ? pk2(ax.z, ax.w) : pk2(ax.x, ax.y), ox));
            const f2 ny = mul2(iy, sub2(q ? pk2(ay.z, ay.w) : pk2(ay.x, ay.y), oy));
            const f2 nz = mul2(iz, sub2(q ? pk2(az.z, az.w) : pk2(az.x, az.y), oz));
            const f2 fx = mul2(ix, sub2(q ? pk2(bx.z, bx.w) : pk2(bx.x, bx.y), ox));
            const f2 fy = mul2(iy, sub2(q ? pk2(by.z, by.w) : pk2(by.x, by.y), oy));
            const f2 fz = mul2(iz, sub2(q ? pk2(bz.z, bz.w) : pk2(bz.x, bz.y), oz));
            float n0, n1, n2, n3, n4, n5, f0, f1, f2_, f3, f4, f5;
            upk2(nx, n0, n1);
            upk2(ny, n2, n3);
            upk2(nz, n4, n5);
            upk2(fx, f0, f1);
            upk2(fy, f2_, f3);
            upk2(fz, f4, f5);
            const float tmin_a = fmaxf(fmaxf(n0, n2), n4), tmin_b = fmaxf(fmaxf(n1, n3), n5);
            float tmax_a, tmax_b;
            upk2(mul2(pk2(fminf(fminf(f0, f2_), f4), fminf(fminf(f1, f3), f5)), slack), tmax_a, tmax_b);
            const int c = h * 4 + q * 2;
            tmin[c] = tmin_a;
            tmin[c + 1] = tmin_b;
            if ((tmin_a <= tmax_a) & (tmin_a <= t) & (tmax_a > 0.0f)) {
                mask |= 1u << c;
            }
            if ((tmin_b <= tmax_b) & (tmin_b <= t) & (tmax_b > 0.0f)) {
                mask |= 2u << c;
            }
        }
    }
    d0 = make_float4(tmin[0], tmin[1], tmin[2], tmin[3]);
    d1 = make_float4(tmin[4], tmin[5], tmin[6], tmin[7]);
    return mask;
}

// ---- cooperative leaf test: 4 lanes per (ray, 8-triangle block) ----------------------------------------------------
// Device triangle copy (k_build_dmtris): block b = 384 bytes, sub-lane k in 0..3 owns bytes [96 k, 96 k + 96) =
//   {nx nx' ny ny'} {nz nz' nw nw'} {ux ux' uy uy'} {uz uz' uw uw'} {vx vx' vy vy'} {vz vz' vw vw'}
// with the unprimed value from triangle k and the primed one from triangle k + 4 of the block: exactly the two
// triangles lane k of the reference's 4-wide test walks (CoreRef.cpp:54-119), adjacent so every multiplication of the
// pair is one FMUL2.
__global__ void k_build_dmtris(const MTri *__restrict__ src, float4 *__restrict__ dst, uint32_t blocks) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = gid >> 2, k = gid & 3u;
    if (b >= blocks) {
        return;
    }
    const MTri &t = src[b];
    float4 *o = dst + size_t(b) * 24 + k * 6;
    o[0] = make_float4(t.n_plane[0][k], t.n_plane[0][k + 4], t.n_plane[1][k], t.n_plane[1][k + 4]);
    o[1] = make_float4(t.n_plane[2][k], t.n_plane[2][k + 4], t.n_plane[3][k], t.n_plane[3][k + 4]);
    o[2] = make_float4(t.u_plane[0][k], t.u_plane[0][k + 4], t.u_plane[1][k], t.u_plane[1][k + 4]);
    o[3] = make_float4(t.u_plane[2][k], t.u_plane[2][k + 4], t.u_plane[3][k], t.u_plane[3][k + 4]);
    o[4] = make_float4(t.v_plane[0][k], t.v_plane[0][k + 4], t.v_plane[1][k], t.v_plane[1][k + 4]);
    o[5] = make_float4(t.v_plane[2][k], t.v_plane[2][k + 4], t.v_plane[3][k], t.v_plane[3][k + 4]);
}

// One lane's share of the 8-wide test: triangle k against t_in, then triangle k + 4 against the lane's updated t.
// Reference IntersectTri(mtri_accel_t) CoreRef.cpp:54-119, same operations in the same order; multiplications two triangles per FMUL2
// (the additions stay scalar: ptxas contracts a packed multiply feeding a packed add into FFMA2 even for mul.rn /
// add.rn.f32x2 and with --fmad=false (CUDA 12.9), which would change the rounding),
// additions scalar (see the note on FFMA2 contraction above).  Returns whether the lane recorded a hit; t/u/v/prim
// then hold the lane's final record (lt, lu, lv, lp of the reference).
RT_DEV bool tri_pair_test(const float4 *__restrict__ q, v3 ro, v3 rd, float t_in, int prim_k, float &t, float &u, float &v,
                          int &prim) {
    const float4 q0 = __ldg(q + 0), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4),
                 q5 = __ldg(q + 5);
    const f2 rdx = bc2(rd.x), rdy = bc2(rd.y), rdz = bc2(rd.z);
    const f2 rox = bc2(ro.x), roy = bc2(ro.y), roz = bc2(ro.z);
    const f2 nx = pk2(q0.x, q0.y), ny = pk2(q0.z, q0.w), nz = pk2(q1.x, q1.y);
    float a[2], b[2], c[2], d[2], det[2], dett[2];
    upk2(mul2(rdx, nx), a[0], a[1]);
    upk2(mul2(rdy, ny), b[0], b[1]);
    upk2(mul2(rdz, nz), c[0], c[1]);
    det[0] = a[0] + b[0] + c[0];
    det[1] = a[1] + b[1] + c[1];
    upk2(mul2(rox, nx), a[0], a[1]);
    upk2(mul2(roy, ny), b[0], b[1]);
    upk2(mul2(roz, nz), c[0], c[1]);
    dett[0] = q1.z - a[0] - b[0] - c[0];
    dett[1] = q1.w - a[1] - b[1] - c[1];
    const f2 det2 = pk2(det[0], det[1]), dett2 = pk2(dett[0], dett[1]);
    float px[2], py[2], pz[2];
    upk2(mul2(det2, rox), a[0], a[1]);
    upk2(mul2(dett2, rdx), b[0], b[1]);
    px[0] = a[0] + b[0];
    px[1] = a[1] + b[1];
    upk2(mul2(det2, roy), a[0], a[1]);
    upk2(mul2(dett2, rdy), b[0], b[1]);
    py[0] = a[0] + b[0];
    py[1] = a[1] + b[1];
    upk2(mul2(det2, roz), a[0], a[1]);
    upk2(mul2(dett2, rdz), b[0], b[1]);
    pz[0] = a[0] + b[0];
    pz[1] = a[1] + b[1];
    const f2 px2 = pk2(px[0], px[1]), py2 = pk2(py[0], py[1]), pz2 = pk2(pz[0], pz[1]);
    float detu[2], detv[2];
    upk2(mul2(px2, pk2(q2.x, q2.y)), a[0], a[1]);
    upk2(mul2(py2, pk2(q2.z, q2.w)), b[0], b[1]);
    upk2(mul2(pz2, pk2(q3.x, q3.y)), c[0], c[1]);
    upk2(mul2(det2, pk2(q3.z, q3.w)), d[0], d[1]);
    detu[0] = a[0] + b[0] + c[0] + d[0];
    detu[1] = a[1] + b[1] + c[1] + d[1];
    upk2(mul2(px2, pk2(q4.x, q4.y)), a[0], a[1]);
    upk2(mul2(py2, pk2(q4.z, q4.w)), b[0], b[1]);
    upk2(mul2(pz2, pk2(q5.x, q5.y)), c[0], c[1]);
    upk2(mul2(det2, pk2(q5.z, q5.w)), d[0], d[1]);
    detv[0] = a[0] + b[0] + c[0] + d[0];
    detv[1] = a[1] + b[1] + c[1] + d[1];
    bool any = false;
    t = t_in;
    u = 0.0f;
    v = 0.0f;
    prim = 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        // sign tests of the reference, in its order; each uses the lane's CURRENT t (updated by triangle k)
        const bool hit = ((__float_as_int(dett[e]) ^ __float_as_int(det[e] * t - dett[e])) >= 0) &&
                         ((__float_as_int(detu[e]) ^ __float_as_int(det[e] - detu[e])) >= 0) &&
                         ((__float_as_int(detv[e]) ^ __float_as_int(det[e] - detu[e] - detv[e])) >= 0);
        const float rdet = 1.0f / (hit ? det[e] : 1.0f);
        if (hit) {
            const int idx = prim_k + 4 * e;
            prim = (det[e] < 0.0f) ? idx : (-idx - 1);
            t = dett[e] * rdet;
            u = detu[e] * rdet;
            v = detv[e] * rdet;
            any = true;
        }
    }
    return any;
}

// ---- per-lane traversal state ---------------------------------------------------------------------------------------
enum : int { LANE_IDLE = 0, LANE_RUN = 1, LANE_FIN = 2 };
enum : uint32_t { TF_IN_BLAS = 1u, TF_RES = 2u, TF_SOLID = 4u };

struct Trav {
    v3 o, d, inv_d;  // the ray in the space of the level being walked
    float t;         // distance of the closest hit so far (hit_data_t::t)
    uint32_t cur;    // node word to visit next, kNoNode = pop
    int sp, base;    // stack size; size at which the current level is drained
    int obj_index;
    uint32_t flags;  // TF_*
};

RT_DEV void trav_begin(Trav &tv, const SceneGeo &sc, TraceSmem &sm, v3 ro, v3 rd) {
    float4 &wo = sm.wro[threadIdx.x], &wd = sm.wrd[threadIdx.x];
    wo.x = ro.x;
    wo.y = ro.y;
    wo.z = ro.z;
    wd.x = rd.x;
    wd.y = rd.y;
    wd.z = rd.z;
    tv.o = ro;
    tv.d = rd;
    tv.inv_d = safe_invert(rd);
    sm.winv[threadIdx.x] = make_float4(tv.inv_d.x, tv.inv_d.y, tv.inv_d.z, 0.0f);
    tv.cur = sc.tlas_root_word;
    tv.sp = 0;
    tv.base = 0;
    tv.obj_index = -1;
    tv.flags = 0;
}

// One pass over the four traversal phases for all lanes of the warp.  `run` = this lane has a traversal in flight.
// Returns true when this lane's traversal has just finished (closest: stack drained; any-hit: also at a solid hit).
template <bool ANY_HIT>
RT_DEV bool trav_step(Trav &tv, const SceneGeo &sc, uint32_t ray_mask, bool run, TraceSmem &sm, TStack &ov,
                      TraverseCounters &cnt) {
    // ---- phase 1: inner node ----
    if (run && int(tv.cur) >= 0 && tv.cur != kNoNode) {
        ++cnt.nodes;
        const char *node = reinterpret_cast<const char *>(sc.dnodes) + size_t(tv.cur) * sizeof(WNode);
        float4 d0, d1;
        sm.cword[0][threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(node + 192));
        sm.cword[1][threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(node + 208));
        uint32_t mask = box8_near_far(node, axis_sel(tv.inv_d), tv.o, tv.inv_d, tv.t, d0, d1);
        const uint32_t *child = reinterpret_cast<const uint32_t *>(&sm.cword[0][threadIdx.x]);
#define RT_CW(i) child[((i) >> 2) * (kTraceThreads * 4) + ((i) & 3)]
        if (mask == 0) {
            tv.cur = kNoNode;
        } else {
            const int i1 = __ffs(mask) - 1;
            mask &= mask - 1;
            if (mask == 0) {
                tv.cur = RT_CW(i1);
            } else {
                sm.dist[0][threadIdx.x] = d0;
                sm.dist[1][threadIdx.x] = d1;
                const float *scr = reinterpret_cast<const float *>(&sm.dist[0][threadIdx.x]);
                // element i of the 8 distances: plane (i >> 2), component (i & 3)
#define RT_SCR(i) scr[((i) >> 2) * (kTraceThreads * 4) + ((i) & 3)]
                const int i2 = __ffs(mask) - 1;
                mask &= mask - 1;
                if (mask == 0) {
                    const float da = RT_SCR(i1), db = RT_SCR(i2);
                    const uint32_t ca = RT_CW(i1), cb = RT_CW(i2);
                    if (da < db) {
                        RT_ST_PUT(sm, ov, tv.sp, cb, db);
                        tv.cur = ca;
                    } else {
                        RT_ST_PUT(sm, ov, tv.sp, ca, da);
                        tv.cur = cb;
                    }
                    ++tv.sp;
                } else {
                    const int i3 = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (mask == 0) {
                        StackEntry e[3] = {{RT_CW(i1), RT_SCR(i1)}, {RT_CW(i2), RT_SCR(i2)},
                                           {RT_CW(i3), RT_SCR(i3)}};
                        sort_top3(e, 3);
                        RT_ST_PUT(sm, ov, tv.sp, e[0].index, e[0].dist);
                        RT_ST_PUT(sm, ov, tv.sp + 1, e[1].index, e[1].dist);
                        tv.sp += 2;
                        tv.cur = e[2].index;
                    } else {
                        const int i4 = __ffs(mask) - 1;
                        mask &= mask - 1;
                        if (mask == 0) {
                            StackEntry e[4] = {{RT_CW(i1), RT_SCR(i1)}, {RT_CW(i2), RT_SCR(i2)},
                                               {RT_CW(i3), RT_SCR(i3)}, {RT_CW(i4), RT_SCR(i4)}};
                            sort_top4(e, 4);
                            RT_ST_PUT(sm, ov, tv.sp, e[0].index, e[0].dist);
                            RT_ST_PUT(sm, ov, tv.sp + 1, e[1].index, e[1].dist);
                            RT_ST_PUT(sm, ov, tv.sp + 2, e[2].index, e[2].dist);
                            tv.sp += 3;
                            tv.cur = e[3].index;
                        } else {
                            // five or more children hit: push all in child order, insertion sort of the whole group
                            // (sort_topN with count = N, CoreRef.cpp:2086-2091), pop the nearest
                            const int first = tv.sp;
                            RT_ST_PUT(sm, ov, tv.sp, RT_CW(i1), RT_SCR(i1));
                            RT_ST_PUT(sm, ov, tv.sp + 1, RT_CW(i2), RT_SCR(i2));
                            RT_ST_PUT(sm, ov, tv.sp + 2, RT_CW(i3), RT_SCR(i3));
                            RT_ST_PUT(sm, ov, tv.sp + 3, RT_CW(i4), RT_SCR(i4));
                            tv.sp += 4;
                            do {
                                const int i = __ffs(mask) - 1;
                                mask &= mask - 1;
                                RT_ST_PUT(sm, ov, tv.sp, RT_CW(i), RT_SCR(i));
                                ++tv.sp;
                            } while (mask != 0);
                            for (int i = first + 1; i < tv.sp; ++i) {
                                const uint2 key = st_get(sm, ov, i);
                                int j = i - 1;
                                while (j >= first) {
                                    const uint2 ej = st_get(sm, ov, j);
                                    if (!(__uint_as_float(ej.y) < __uint_as_float(key.y))) {
                                        break;
                                    }
                                    RT_ST_PUT(sm, ov, j + 1, ej.x, __uint_as_float(ej.y));
                                    --j;
                                }
                                RT_ST_PUT(sm, ov, j + 1, key.x, __uint_as_float(key.y));
                            }
                            tv.cur = st_get(sm, ov, --tv.sp).x;
                        }
                    }
                }
#undef RT_SCR
#undef RT_CW
            }
        }
    }
    __syncwarp();
    // ---- phase 2: BLAS leaves, cooperatively: the lanes holding a leaf queue their ray in shared memory, then every
    // group of 4 lanes tests one (ray, block): sub-lane k walks triangles k and k + 4 like lane k of the reference ----
    bool finished = false;
    const bool at_leaf = run && int(tv.cur) < 0 && (tv.flags & TF_IN_BLAS);
    const uint32_t leafmask = __ballot_sync(0xffffffffu, at_leaf);
    if (leafmask != 0) {
        const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
        float4 *slot_o = &sm.dist[0][wbase], *slot_d = &sm.dist[1][wbase];
        float4 *slot_r = reinterpret_cast<float4 *>(&sm.cword[0][wbase]);
        const int rank = __popc(leafmask & ((1u << lane) - 1u));
        if (at_leaf) {
            ++cnt.leaves;
            slot_o[rank] = make_float4(tv.o.x, tv.o.y, tv.o.z, tv.t);
            slot_d[rank] = make_float4(tv.d.x, tv.d.y, tv.d.z, __uint_as_float((tv.cur & kLeafFirstBits) >> 3));
            slot_r[rank] = make_float4(0.0f, 0.0f, -1.0f, 0.0f);
        }
        __syncwarp();
        const int n = __popc(leafmask);
        const int g = lane >> 2, k = lane & 3;
        for (int r0 = 0; r0 < n; r0 += 8) {
            const int r = r0 + g;
            const bool valid = r < n;
            const float4 so = slot_o[valid ? r : 0], sd = slot_d[valid ? r : 0];
            const uint32_t block = __float_as_uint(sd.w);
            const float4 *q = reinterpret_cast<const float4 *>(sc.dmtris) + size_t(block) * 24 + k * 6;
            float lt, lu, lv;
            int lp;
            const bool any = tri_pair_test(q, v3{so.x, so.y, so.z}, v3{sd.x, sd.y, sd.z}, so.w, int(block) * 8 + k, lt, lu,
                                           lv, lp);
            // min over the 4 lanes in the reference's order, lowest lane holding it wins
            const int qb = lane & ~3;
            const float t0 = __shfl_sync(0xffffffffu, lt, qb + 0), t1 = __shfl_sync(0xffffffffu, lt, qb + 1),
                        t2 = __shfl_sync(0xffffffffu, lt, qb + 2), t3 = __shfl_sync(0xffffffffu, lt, qb + 3);
            const float min_t = fminf(t0, fminf(t1, fminf(t2, t3)));
            const uint32_t wins = (__ballot_sync(0xffffffffu, valid && any && lt == min_t) >> qb) & 0xfu;
            if (wins != 0 && k == __ffs(wins) - 1) {
                slot_r[r] = make_float4(lt, lu, lv, __int_as_float(lp));
            }
        }
        __syncwarp();
        if (at_leaf) {
            const float4 res = slot_r[rank];
            const uint32_t more = (tv.cur >> kLeafBlocksShift) & 15u;
            // a leaf spanning several blocks (never produced by the builders seen so far) walks them one phase at a time
            tv.cur = more ? (kLeafBit | ((more - 1u) << kLeafBlocksShift) | ((tv.cur & kLeafFirstBits) + 8u)) : kNoNode;
            if (res.z >= 0.0f) {
                const int prim = __float_as_int(res.w);
                tv.t = res.x;
                sm.hit[threadIdx.x] = make_float4(res.y, res.z, res.w, __int_as_float(tv.obj_index));
                tv.flags |= TF_RES;
                if (ANY_HIT) {
                    const bool backfacing = prim < 0;
                    const uint32_t slot = backfacing ? uint32_t(-prim - 1) : uint32_t(prim);
                    const TriMat tm = sc.tri_materials[__ldg(&sc.tri_indices[slot])];
                    if ((!backfacing && (tm.front_mi & kMatSolidBit)) || (backfacing && (tm.back_mi & kMatSolidBit))) {
                        tv.flags |= TF_SOLID;
                        tv.cur = kNoNode;
                        finished = true;
                    }
                }
            }
        }
    }
    __syncwarp();
    // ---- phase 3: TLAS leaf = one mesh instance (CoreRef.cpp:2098-2119, TransformRay :2789-2798) ----
    if (run && int(tv.cur) < 0 && !(tv.flags & TF_IN_BLAS) && !finished) {
        const uint32_t mi_index = tv.cur & kLeafFirstBits;
        const MeshInstance *__restrict__ mi = sc.instances + mi_index;
        if ((__ldg(&mi->ray_visibility) & ray_mask) == 0) {
            tv.cur = kNoNode;
        } else {
            const float *__restrict__ m = mi->inv_xform;
            const float4 wo = sm.wro[threadIdx.x], wd = sm.wrd[threadIdx.x];
            const v3 ro = v3{wo.x, wo.y, wo.z}, rd = v3{wd.x, wd.y, wd.z};
            tv.o = v3{m[0] * ro.x + m[4] * ro.y + m[8] * ro.z + m[12], m[1] * ro.x + m[5] * ro.y + m[9] * ro.z + m[13],
                      m[2] * ro.x + m[6] * ro.y + m[10] * ro.z + m[14]};
            tv.d = v3{m[0] * rd.x + m[4] * rd.y + m[8] * rd.z, m[1] * rd.x + m[5] * rd.y + m[9] * rd.z,
                      m[2] * rd.x + m[6] * rd.y + m[10] * rd.z};
            tv.inv_d = safe_invert(tv.d);
            tv.flags |= TF_IN_BLAS;
            tv.obj_index = int(mi_index);
            tv.base = tv.sp;
            tv.cur = __ldg(&sc.blas_roots[mi_index]); // the reference pushes the BLAS root with dist 0 and pops it at once
        }
    }
    __syncwarp();
    // ---- phase 4: pop (cull entries farther than the current hit) ----
    if (run && !finished) {
        while (tv.cur == kNoNode) {
            if (tv.sp == tv.base) {
                if (!(tv.flags & TF_IN_BLAS)) {
                    finished = true;
                    break;
                }
                // BLAS drained: back to the TLAS level with the world-space ray
                tv.flags &= ~TF_IN_BLAS;
                tv.base = 0;
                const float4 wo = sm.wro[threadIdx.x], wd = sm.wrd[threadIdx.x], wi = sm.winv[threadIdx.x];
                tv.o = v3{wo.x, wo.y, wo.z};
                tv.d = v3{wd.x, wd.y, wd.z};
                tv.inv_d = v3{wi.x, wi.y, wi.z};
                continue;
            }
            const uint2 e = st_get(sm, ov, --tv.sp);
            if (!(__uint_as_float(e.y) > tv.t)) {
                tv.cur = e.x;
            }
        }
    }
    __syncwarp();
    return finished;
}

// the hit record of a finished traversal; the primitive index indirection is resolved also for misses, like the
// reference (CoreRef.cpp:2125-2130)
RT_DEV Hit trav_result(const Trav &tv, const SceneGeo &sc, const TraceSmem &sm, bool resolve) {
    const float4 h = sm.hit[threadIdx.x];
    Hit r;
    r.t = tv.t;
    r.u = h.x;
    r.v = h.y;
    r.prim = __float_as_int(h.z);
    r.obj = __float_as_int(h.w);
    if (resolve) {
        if (r.prim < 0) {
            r.prim = -int(__ldg(&sc.tri_indices[-r.prim - 1])) - 1;
        } else {
            r.prim = int(__ldg(&sc.tri_indices[r.prim]));
        }
    }
    return r;
}

RT_DEV void trav_set_hit(Trav &tv, TraceSmem &sm, const Hit &h) {
    tv.t = h.t;
    sm.hit[threadIdx.x] = make_float4(h.u, h.v, __int_as_float(h.prim), __int_as_float(h.obj));
}

// ---- work distribution: lanes take rays one at a time ---------------------------------------------------------------
// The first `pk` lanes of every warp start on a static slice (no atomic); afterwards idle lanes are refilled from the
// list's queue head, one atomic per refill for all lanes of the warp that need a ray.
struct LaneQueue {
    uint32_t count, static_end;
    bool exhausted; // warp-uniform
};

RT_DEV uint32_t lane_queue_init(LaneQueue &q, uint32_t count, int lane, bool &has) {
    q.count = count;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t pk = 32;
    while (pk > 4 && uint64_t(pk / 2) * warps >= count) {
        pk >>= 1;
    }
    q.static_end = warps * pk;
    q.exhausted = q.static_end >= count;
    const uint32_t i = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * pk + uint32_t(lane);
    has = (uint32_t(lane) < pk) && i < count;
    return i;
}

// `want` = warp ballot of the lanes that need a ray.  Returns this lane's new ray index (valid where `has`).
RT_DEV uint32_t lane_queue_take(LaneQueue &q, uint32_t *head, uint32_t want, int lane, bool &has) {
    has = false;
    if (q.exhausted || want == 0) {
        return 0;
    }
    const int leader = __ffs(want) - 1;
    const uint32_t n = __popc(want);
    uint32_t b = 0;
    if (lane == leader) {
        b = atomicAdd(head, n);
    }
    b = q.static_end + __shfl_sync(0xffffffffu, b, leader);
    if (b + n >= q.count) {
        q.exhausted = true;
    }
    const uint32_t i = b + __popc(want & ((1u << lane) - 1u));
    has = ((want >> lane) & 1u) && i < q.count;
    return i;
}

RT_DEV void flush_counters(const KParams &p, TraverseCounters cnt, int lane) {
    uint32_t n = cnt.nodes, l = cnt.leaves;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, off);
        l += __shfl_xor_sync(0xffffffffu, l, off);
    }
    if (lane == 0 && (n | l)) {
        atomicAdd(&p.totals[TOT_NODES], (unsigned long long)n);
        atomicAdd(&p.totals[TOT_LEAVES], (unsigned long long)l);
    }
}

// ---- TraceRays: IntersectScene (CoreRef.cpp:3041-3158) [+ IntersectAreaLights :3616-3860] ------------------------
// INIT_HITS: secondary lists start from the default "no intersection" record (RendererCPU.h:532-535) built in
// registers instead of a memset pass + 20 B/ray read.  `fin_min`: lanes whose traversal has finished wait until that
// many of the warp have (or nothing else is running) before the per-ray epilogue + refill code is issued for them.
// Per-ray words in shared memory: wro.w = t the traversal in flight started with (t_val), wrd.w = bits(depth).
template <bool TRACE_LIGHTS, bool INIT_HITS>
__global__ void __launch_bounds__(RT_TRACE_THREADS, RT_TRACE_BLOCKS)
    k_trace_closest(KParams p, RayBuf rays, HitBuf hits, int bounce, int fin_min) {
    __shared__ TraceSmem sm;
    TStack ov;
    const uint32_t count = p.counters[CNT_RAYS + bounce];
    uint32_t *head = &p.counters[CNT_HEAD_TRACE + bounce];
    const int lane = threadIdx.x & 31;
    const SceneGeo &sc = p.sc.geo;
    TraverseCounters cnt{0, 0};
    LaneQueue q;
    Trav tv;
    uint32_t ray_mask = 0;
    int state = LANE_IDLE;
    bool fresh;
    sm.idx[threadIdx.x] = lane_queue_init(q, count, lane, fresh);

    while (true) {
        // ---- epilogue of finished traversals + refill ----
        const uint32_t fin = __ballot_sync(0xffffffffu, state == LANE_FIN);
        const uint32_t running = __ballot_sync(0xffffffffu, state == LANE_RUN);
        const bool do_fin = fin != 0 && (int(__popc(fin)) >= fin_min || running == 0);
        if (do_fin && state == LANE_FIN) {
            const uint32_t i = sm.idx[threadIdx.x];
            Hit inter = trav_result(tv, sc, sm, true);
            const float4 wo = sm.wro[threadIdx.x], wd = sm.wrd[threadIdx.x];
            const v3 ro = v3{wo.x, wo.y, wo.z}, rd = v3{wd.x, wd.y, wd.z};
            uint32_t depth = __float_as_uint(wd.w);
            bool again = false;
            if (tv.flags & TF_RES) {
                const bool is_backfacing = (inter.prim < 0);
                const uint32_t tri_index = is_backfacing ? uint32_t(-inter.prim - 1) : uint32_t(inter.prim);
                const TriMat tm = sc.tri_materials[tri_index];
                const bool solid = (!is_backfacing && (tm.front_mi & kMatSolidBit)) || (is_backfacing && (tm.back_mi & kMatSolidBit));
                if (!solid) {
                    // the transparency loop of IntersectScene: stochastic Mix resolve, Transparent -> continue the ray
                    const Material *mat = is_backfacing ? &p.sc.surf.materials[tm.back_mi & kMatIndexBits]
                                                        : &p.sc.surf.materials[tm.front_mi & kMatIndexBits];
                    const uint32_t xy = rays.xy_depth[i].x;
                    const uint32_t rand_dim = kRandDimBase + total_depth(depth) * kRandDimBounce;
                    const uint32_t rand_hash = hash_combine(hash_u32(xy), p.rand_seed);
                    const v2 mix_term_rand = rand2d(rand_dim + kRandDimBsdfPick, rand_hash, p.iteration - 1, p.sc.rand_seq);
                    float trans_r = mix_term_rand.x;
                    // alpha-textured Mix nodes (CoreRef.cpp:3088-3115): uvs at the hit + the bounce's texture jitter
                    v2 uvs = v2{0.0f, 0.0f}, tex_rand = v2{0.0f, 0.0f};
                    if (p.sc.tex.descs != nullptr && mat->type == NODE_MIX) {
                        const Vertex &v1 = p.sc.surf.vertices[p.sc.surf.vtx_indices[tri_index * 3 + 0]];
                        const Vertex &v2_ = p.sc.surf.vertices[p.sc.surf.vtx_indices[tri_index * 3 + 1]];
                        const Vertex &v3_ = p.sc.surf.vertices[p.sc.surf.vtx_indices[tri_index * 3 + 2]];
                        const float w = 1.0f - inter.u - inter.v;
                        uvs = v2{v1.t[0] * w + v2_.t[0] * inter.u + v3_.t[0] * inter.v,
                                 v1.t[1] * w + v2_.t[1] * inter.u + v3_.t[1] * inter.v};
                        tex_rand = rand2d(rand_dim + kRandDimTex, rand_hash, p.iteration - 1, p.sc.rand_seq);
                    }
                    while (mat->type == NODE_MIX) {
                        float mix_val = mat->tangent_rotation_or_strength;
                        const uint32_t mix_texture = mat->textures[kTexBase];
                        if (mix_texture != kTexInvalid) {
                            mix_val *= tex_sample_color(p.sc.tex, mix_texture, uvs, 0, tex_rand, true).x;
                        }
                        if (trans_r > mix_val) {
                            mat = &p.sc.surf.materials[mat->textures[kMixMat1]];
                            trans_r = safe_div_pos(trans_r - mix_val, 1.0f - mix_val);
                        } else {
                            mat = &p.sc.surf.materials[mat->textures[kMixMat2]];
                            trans_r = safe_div_pos(trans_r, mix_val);
                        }
                    }
                    if (mat->type == NODE_TRANSPARENT) {
                        // throughput and depth of the ray live in its record: this kernel is the only one touching it
                        float4 c = rays.c_pdf[i];
                        const bool can_terminate_path = transp_depth(depth) > p.ps.min_transp_depth;
                        const float lum_ = fmaxf(c.x, fmaxf(c.y, c.z));
                        const float pr = mix_term_rand.y;
                        const float qq = can_terminate_path ? fmaxf(0.05f, 1.0f - lum_) : 0.0f;
                        if (pr < qq || lum_ == 0.0f || transp_depth(depth) + 1 >= p.ps.max_transp_depth) {
                            c.x = c.y = c.z = 0.0f;
                        } else {
                            c.x *= mat->base_color[0] / (1.0f - qq);
                            c.y *= mat->base_color[1] / (1.0f - qq);
                            c.z *= mat->base_color[2] / (1.0f - qq);
                            const float t = inter.t + kHitBias;
                            inter.v = -1.0f;
                            inter.t = wo.w - inter.t; // t_val - t
                            depth += pack_depth(0, 0, 0, 1);
                            trav_begin(tv, sc, sm, ro + rd * t, rd);
                            trav_set_hit(tv, sm, inter);
                            sm.wro[threadIdx.x].w = inter.t;
                            sm.wrd[threadIdx.x].w = __uint_as_float(depth);
                            again = true;
                        }
                        rays.c_pdf[i] = c;
                        rays.xy_depth[i] = make_uint2(xy, depth);
                    }
                }
            }
            if (again) {
                state = LANE_RUN;
            } else {
                const float4 a = rays.o_cw[i];
                const v3 r_o = v3{a.x, a.y, a.z};
                inter.t += length(r_o - ro);
                if (TRACE_LIGHTS) {
                    if (p.sc.lights.visible_lights_count != 0) {
                        LightStackEntry lst[kMaxStack];
                        intersect_area_lights(p.sc.lights, r_o, rd, ray_mask, inter, lst);
                    }
                }
                store_hit(hits, i, inter);
                state = LANE_IDLE;
            }
        }
        if (do_fin || __any_sync(0xffffffffu, fresh)) { // warp-uniform
            // refill: lanes without a ray (just finished) take the next one from the queue
            if (do_fin) {
                const uint32_t want = __ballot_sync(0xffffffffu, state == LANE_IDLE && !fresh);
                bool has;
                const uint32_t ni = lane_queue_take(q, head, want, lane, has);
                if (has) {
                    sm.idx[threadIdx.x] = ni;
                    fresh = true;
                }
            }
            if (fresh) {
                const uint32_t i = sm.idx[threadIdx.x];
                const float4 a = rays.o_cw[i], dd = rays.d_cs[i];
                const uint32_t depth = rays.xy_depth[i].y;
                trav_begin(tv, sc, sm, v3{a.x, a.y, a.z}, v3{dd.x, dd.y, dd.z});
                Hit h0;
                if (INIT_HITS) {
                    h0.obj = -1;
                    h0.prim = -1;
                    h0.t = kMaxDist;
                    h0.u = 0.0f;
                    h0.v = -1.0f;
                } else {
                    h0 = load_hit(hits, i);
                }
                trav_set_hit(tv, sm, h0);
                sm.wro[threadIdx.x].w = h0.t;
                sm.wrd[threadIdx.x].w = __uint_as_float(depth);
                ray_mask = 1u << ray_type(depth);
                state = LANE_RUN;
                fresh = false;
            }
        }
        __syncwarp();
        if (__ballot_sync(0xffffffffu, state != LANE_IDLE) == 0) {
            break;
        }
        if (trav_step<false>(tv, sc, ray_mask, state == LANE_RUN, sm, ov, cnt)) {
            state = LANE_FIN;
        }
    }
    flush_counters(p, cnt, lane);
}

// blocker lights, clamp and the add into the radiance plane (<= 1 shadow ray per pixel per stage: plain RMW)
RT_DEV void shadow_finish(const KParams &p, const ShadowBuf &srays, uint32_t i, v3 rc, float limit) {
    const float4 cx = srays.c_xy[i];
    if (p.sc.lights.blocker_lights_count != 0) {
        const float4 od = srays.o_depth[i], dd = srays.d_dist[i];
        StackEntry lst[kMaxStack];
        rc *= intersect_area_lights_shadow(p.sc.lights, v3{od.x, od.y, od.z}, v3{dd.x, dd.y, dd.z}, dd.w, lst);
    }
    const float sum = ((rc.x + rc.y) + rc.z) + 0.0f;
    if (sum > limit) {
        rc *= (limit / sum);
    }
    const uint32_t xy = __float_as_uint(cx.w);
    const int x = int((xy >> 16) & 0xffff), y = int(xy & 0xffff);
    float4 o = p.fb.temp[y * p.fb.w + x];
    o.x += rc.x;
    o.y += rc.y;
    o.z += rc.z;
    o.w += 0.0f;
    p.fb.temp[y * p.fb.w + x] = o;
}

// ---- TraceShadowRays (CoreRef.cpp:4856-4882) + IntersectScene(shadow) (:3160-3262) -------------------------------
// Per-ray words in shared memory: wro.w = remaining distance, wrd.w = bits(transparency depth), aux = throughput.
__global__ void __launch_bounds__(RT_TRACE_THREADS, RT_TRACE_BLOCKS)
    k_trace_shadow(KParams p, ShadowBuf srays, int stage, float limit, int fin_min) {
    __shared__ TraceSmem sm;
    TStack ov;
    const uint32_t count = p.counters[CNT_SHADOW + stage];
    uint32_t *head = &p.counters[CNT_HEAD_SHADOW + stage];
    const int lane = threadIdx.x & 31;
    const SceneGeo &sc = p.sc.geo;
    TraverseCounters cnt{0, 0};
    LaneQueue q;
    Trav tv;
    int state = LANE_IDLE;
    bool fresh;
    sm.idx[threadIdx.x] = lane_queue_init(q, count, lane, fresh);

    while (true) {
        const uint32_t fin = __ballot_sync(0xffffffffu, state == LANE_FIN);
        const uint32_t running = __ballot_sync(0xffffffffu, state == LANE_RUN);
        const bool do_fin = fin != 0 && (int(__popc(fin)) >= fin_min || running == 0);
        if (do_fin && state == LANE_FIN) {
            // body of the `while (dist > HIT_BIAS)` loop of IntersectScene(shadow) after the traversal
            const uint32_t i = sm.idx[threadIdx.x];
            const bool solid_hit = (tv.flags & TF_SOLID) != 0;
            const float4 wo = sm.wro[threadIdx.x], wd = sm.wrd[threadIdx.x], ax = sm.aux[threadIdx.x];
            v3 rc = v3{ax.x, ax.y, ax.z};
            float dist = wo.w;
            int depth = __float_as_int(wd.w);
            bool again = false;
            if (solid_hit || depth > p.ps.max_transp_depth) {
                rc = v3{0.0f, 0.0f, 0.0f};
            } else {
                const Hit inter = trav_result(tv, sc, sm, true);
                if (!(inter.v < 0.0f)) {
                    const bool is_backfacing = (inter.prim < 0);
                    const uint32_t tri_index = is_backfacing ? uint32_t(-inter.prim - 1) : uint32_t(inter.prim);
                    const TriMat tm = sc.tri_materials[tri_index];
                    const uint32_t mat_index = is_backfacing ? (tm.back_mi & kMatIndexBits) : (tm.front_mi & kMatIndexBits);
                    // transparency throughput of the (possibly mixed) material
                    uint32_t mstack[16];
                    float wstack[16];
                    int ms = 0;
                    mstack[ms] = mat_index;
                    wstack[ms++] = 1.0f;
                    v3 throughput = v3{0.0f, 0.0f, 0.0f};
                    // alpha-textured Mix nodes (CoreRef.cpp:3203-3240)
                    v2 sh_uvs = v2{0.0f, 0.0f}, tex_rand = v2{0.0f, 0.0f};
                    if (p.sc.tex.descs != nullptr) {
                        const Vertex &v1 = p.sc.surf.vertices[p.sc.surf.vtx_indices[tri_index * 3 + 0]];
                        const Vertex &v2_ = p.sc.surf.vertices[p.sc.surf.vtx_indices[tri_index * 3 + 1]];
                        const Vertex &v3_ = p.sc.surf.vertices[p.sc.surf.vtx_indices[tri_index * 3 + 2]];
                        const float w = 1.0f - inter.u - inter.v;
                        sh_uvs = v2{v1.t[0] * w + v2_.t[0] * inter.u + v3_.t[0] * inter.v,
                                    v1.t[1] * w + v2_.t[1] * inter.u + v3_.t[1] * inter.v};
                        // rand_dim advances with the transparency depth (both start from the ray's packed depth word)
                        const uint32_t d0 = __float_as_uint(srays.o_depth[i].w);
                        const uint32_t rand_dim = kRandDimBase + uint32_t(total_depth(d0) + depth - transp_depth(d0)) * kRandDimBounce;
                        const uint32_t xy = __float_as_uint(srays.c_xy[i].w);
                        tex_rand = rand2d(rand_dim + kRandDimTex, hash_combine(hash_u32(xy), p.rand_seed), p.iteration - 1,
                                          p.sc.rand_seq);
                    }
                    while (ms--) {
                        const Material *mat = &p.sc.surf.materials[mstack[ms]];
                        const float weight = wstack[ms];
                        if (mat->type == NODE_MIX) {
                            float mix_val = mat->tangent_rotation_or_strength;
                            const uint32_t mix_texture = mat->textures[kTexBase];
                            if (mix_texture != kTexInvalid) {
                                mix_val *= tex_sample_color(p.sc.tex, mix_texture, sh_uvs, 0, tex_rand, true).x;
                            }
                            mstack[ms] = mat->textures[kMixMat1];
                            wstack[ms++] = weight * (1.0f - mix_val);
                            mstack[ms] = mat->textures[kMixMat2];
                            wstack[ms++] = weight * mix_val;
                        } else if (mat->type == NODE_TRANSPARENT) {
                            throughput += weight * mk3(mat->base_color);
                        }
                    }
                    rc *= throughput;
                    if (!(lum(rc) < kFltEps)) {
                        const float t = inter.t + kHitBias;
                        const v3 ro = v3{wo.x, wo.y, wo.z}, rd = v3{wd.x, wd.y, wd.z};
                        dist -= t;
                        ++depth;
                        if (dist > kHitBias) {
                            trav_begin(tv, sc, sm, ro + rd * t, rd);
                            Hit h0;
                            h0.obj = -1;
                            h0.prim = -1;
                            h0.t = dist;
                            h0.u = 0.0f;
                            h0.v = -1.0f;
                            trav_set_hit(tv, sm, h0);
                            sm.wro[threadIdx.x].w = dist;
                            sm.wrd[threadIdx.x].w = __int_as_float(depth);
                            sm.aux[threadIdx.x] = make_float4(rc.x, rc.y, rc.z, 0.0f);
                            again = true;
                        }
                    }
                }
            }
            if (again) {
                state = LANE_RUN;
            } else {
                shadow_finish(p, srays, i, rc, limit);
                state = LANE_IDLE;
            }
        }
        if (do_fin || __any_sync(0xffffffffu, fresh)) { // warp-uniform
            if (do_fin) {
                const uint32_t want = __ballot_sync(0xffffffffu, state == LANE_IDLE && !fresh);
                bool has;
                const uint32_t ni = lane_queue_take(q, head, want, lane, has);
                if (has) {
                    sm.idx[threadIdx.x] = ni;
                    fresh = true;
                }
            }
            if (fresh) {
                const uint32_t i = sm.idx[threadIdx.x];
                const ShadowRayD r = load_shadow(srays, i);
                const float dist = r.dist > 0.0f ? r.dist : kMaxDist;
                fresh = false;
                if (dist > kHitBias) {
                    trav_begin(tv, sc, sm, r.o, r.d);
                    Hit h0;
                    h0.obj = -1;
                    h0.prim = -1;
                    h0.t = dist;
                    h0.u = 0.0f;
                    h0.v = -1.0f;
                    trav_set_hit(tv, sm, h0);
                    sm.wro[threadIdx.x].w = dist;
                    sm.wrd[threadIdx.x].w = __int_as_float(transp_depth(r.depth));
                    sm.aux[threadIdx.x] = make_float4(r.c.x, r.c.y, r.c.z, 0.0f);
                    state = LANE_RUN;
                } else {
                    // the `while (dist > HIT_BIAS)` loop does not run at all: only the blocker lights and the add
                    shadow_finish(p, srays, i, r.c, limit);
                    state = LANE_IDLE;
                }
            }
        }
        __syncwarp();
        if (__ballot_sync(0xffffffffu, state != LANE_IDLE) == 0) {
            break;
        }
        if (trav_step<true>(tv, sc, (1u << RAY_SHADOW), state == LANE_RUN, sm, ov, cnt)) {
            state = LANE_FIN;
        }
    }
    flush_counters(p, cnt, lane);
}

} // namespace rt
