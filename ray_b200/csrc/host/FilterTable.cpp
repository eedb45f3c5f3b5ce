// FilterTable.cpp -- 1024-entry inverse-CDF table of the pixel reconstruction filter.
//
// Role in the reference: Cpu::Renderer<P>::UpdateFilterTable (internal/RendererCPU.h:1234-1258) over Ray::CDFInverted
// (internal/CDFUtils.{h,cpp}) with the filter shapes of internal/Core.h:316-325.  GeneratePrimaryRays looks the jitter
// of non-Box filters up in this table (internal/CoreRef.cpp:1453-1468).  Same numerical procedure, written for this
// library: tabulate |f| on [0, width/2], accumulate + normalise a CDF, invert it on a symmetric grid.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../rt_types.h"
#include "RendererCuda.h"

namespace RayB200 {
namespace Cuda {

namespace {
constexpr float PI = 3.141592653589793238463f;
float filter_eval(uint32_t filter, float v, float width) {
    switch (filter) {
    case RS_FILTER_GAUSSIAN: {
        v *= 6.0f / width;
        return expf(-2.0f * v * v);
    }
    case RS_FILTER_BLACKMAN_HARRIS: {
        v = 2.0f * PI * (v / width + 0.5f);
        return 0.35875f - 0.48829f * cosf(v) + 0.14128f * cosf(2.0f * v) - 0.01168f * cosf(3.0f * v);
    }
    default: return 1.0f;
    }
}
} // namespace

std::vector<float> GenerateFilterTable(const uint32_t filter, float filter_width) {
    const int res = rt::kFilterTableSize;
    switch (filter) {
    case RS_FILTER_GAUSSIAN: filter_width *= 3.0f; break;
    case RS_FILTER_BLACKMAN_HARRIS: filter_width *= 2.0f; break;
    default: filter_width = 1.0f; break;
    }
    const float from = 0.0f, to = filter_width * 0.5f;
    // CDF over res-1 cells (res entries)
    const int cells = res - 1;
    std::vector<float> cdf(cells + 1);
    cdf[0] = 0.0f;
    const float range = to - from;
    for (int i = 0; i < cells; ++i) {
        const float x = from + range * float(i) / float(cells - 1);
        cdf[i + 1] = cdf[i] + std::fabs(filter_eval(filter, x, filter_width));
    }
    const float fac = (cdf[cells] == 0.0f) ? 0.0f : 1.0f / cdf[cells];
    for (float &c : cdf) {
        c *= fac;
    }
    cdf[cells] = 1.0f;
    // symmetric inversion
    std::vector<float> inv(res);
    const int cdf_size = int(cdf.size());
    const int half = (res - 1) / 2;
    for (int i = 0; i <= half; ++i) {
        const float x = float(i) / float(half);
        int index = int(std::upper_bound(cdf.begin(), cdf.end(), x) - cdf.begin());
        float t;
        if (index < cdf_size - 1) {
            t = (x - cdf[index]) / (cdf[index + 1] - cdf[index]);
        } else {
            t = 0.0f;
            index = cdf_size - 1;
        }
        const float y = ((index + t) / (res - 1)) * 2.0f * range;
        inv[half + i] = 0.5f * (1.0f + y);
        inv[half - i] = 0.5f * (1.0f - y);
    }
    return inv;
}

} // namespace Cuda
} // namespace RayB200
