#pragma once

#define ENABLE_REF_IMPL
#define ENABLE_SIMD_IMPL
/* #undef ENABLE_VK_IMPL */
/* #undef ENABLE_DX_IMPL */
/* #undef ENABLE_GPU_DEBUG */
/* #undef ENABLE_PIX */
