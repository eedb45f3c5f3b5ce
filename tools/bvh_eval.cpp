// Dev tool (not product): count BVH8 node / leaf visits of an ordered closest-hit walk over an rc_scene_view.
// Same visit policy as the kernels (front-to-back by box entry distance, cull against the running t); float math is
// NOT the bit-exact one -- this only ranks acceleration structures.   g++ -O2 -shared -fPIC -I../include
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include "ray_cuda.h"

namespace {
struct WNode { float mn[3][8], mx[3][8]; uint32_t child[8]; };
struct MTri { float n[4][8], u[4][8], v[4][8]; };
struct MeshInstance { uint32_t mesh_index, node_index, lights_index, ray_visibility; float xform[16], inv_xform[16]; };
constexpr uint32_t LEAF = 0x80000000u, EMPTY = 0x7fffffffu;
struct Entry { uint32_t idx; float d; };
}

extern "C" void bvh_eval(const rc_scene_view *v, const float *ro, const float *rd, const float *tmax, int n, int any_hit,
                         uint64_t *out_nodes, uint64_t *out_leaves, float *out_t) {
    const WNode *nodes = (const WNode *)v->wnodes.ptr;
    const MTri *mtris = (const MTri *)v->mtris.ptr;
    const MeshInstance *mis = (const MeshInstance *)v->mesh_instances.ptr;
    uint64_t nn = 0, nl = 0;
    for (int r = 0; r < n; ++r) {
        const float *o = ro + 3 * r, *d = rd + 3 * r;
        float t = tmax ? tmax[r] : 3.4e30f;
        // single instance scenes only need the world ray; instances are handled by transforming on entry
        Entry st[256]; int sp = 0;
        st[sp++] = {v->tlas_root, 0.0f};
        float co[3] = {o[0], o[1], o[2]}, cd[3] = {d[0], d[1], d[2]};
        int blas_base = -1; bool hit = false;
        while (sp) {
            if (blas_base >= 0 && sp == blas_base) { blas_base = -1; memcpy(co, o, 12); memcpy(cd, d, 12); continue; }
            Entry e = st[--sp];
            if (e.d > t) continue;
            const WNode &nd = nodes[e.idx];
            if (nd.child[0] & LEAF) {
                if (blas_base < 0) { // instance
                    const MeshInstance &mi = mis[nd.child[0] & ~LEAF];
                    const float *m = mi.inv_xform;
                    for (int k = 0; k < 3; ++k) {
                        co[k] = m[k] * o[0] + m[4 + k] * o[1] + m[8 + k] * o[2] + m[12 + k];
                        cd[k] = m[k] * d[0] + m[4 + k] * d[1] + m[8 + k] * d[2];
                    }
                    blas_base = sp;
                    st[sp++] = {mi.node_index, 0.0f};
                    continue;
                }
                ++nl;
                const uint32_t start = nd.child[0] & ~LEAF, cnt = nd.child[1];
                for (uint32_t b = start / 8; b < (start + cnt + 7) / 8; ++b) {
                    const MTri &m = mtris[b];
                    for (int k = 0; k < 8; ++k) {
                        const float det = cd[0] * m.n[0][k] + cd[1] * m.n[1][k] + cd[2] * m.n[2][k];
                        const float dett = m.n[3][k] - co[0] * m.n[0][k] - co[1] * m.n[1][k] - co[2] * m.n[2][k];
                        if (det == 0.0f) continue;
                        const float tt = dett / det;
                        if (!(tt > 0.0f) || tt > t) continue;
                        const float p[3] = {det * co[0] + dett * cd[0], det * co[1] + dett * cd[1], det * co[2] + dett * cd[2]};
                        const float du = p[0] * m.u[0][k] + p[1] * m.u[1][k] + p[2] * m.u[2][k] + det * m.u[3][k];
                        const float dv = p[0] * m.v[0][k] + p[1] * m.v[1][k] + p[2] * m.v[2][k] + det * m.v[3][k];
                        const float uu = du / det, vv = dv / det;
                        if (uu < 0.0f || vv < 0.0f || uu + vv > 1.0f) continue;
                        t = tt; hit = true;
                    }
                }
                if (any_hit && hit) break;
                continue;
            }
            ++nn;
            Entry kids[8]; int nk = 0;
            for (int k = 0; k < 8; ++k) {
                if (nd.child[k] == EMPTY) continue;
                float tmin = 0.0f, tmx = t; bool ok = true;
                for (int a = 0; a < 3; ++a) {
                    const float inv = 1.0f / (std::fabs(cd[a]) > 1e-7f ? cd[a] : std::copysign(1e-7f, cd[a]));
                    float lo = (nd.mn[a][k] - co[a]) * inv, hi = (nd.mx[a][k] - co[a]) * inv;
                    if (lo > hi) std::swap(lo, hi);
                    tmin = std::max(tmin, lo); tmx = std::min(tmx, hi * 1.00000024f);
                }
                ok = tmin <= tmx;
                if (ok) kids[nk++] = {nd.child[k], tmin};
            }
            std::sort(kids, kids + nk, [](const Entry &a, const Entry &b) { return a.d > b.d; });
            for (int k = 0; k < nk; ++k) st[sp++] = kids[k];
        }
        if (out_t) out_t[r] = hit ? t : -1.0f;
    }
    *out_nodes = nn; *out_leaves = nl;
}
