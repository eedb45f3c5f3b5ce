"""GPU parity: every stage of the CUDA hot path against the reference's own Ref:: functions (oracle/_ref), fed the
same inputs on byte-identical scene arrays, through the C-ABI stage entry points.  Bar: BIT-EXACT records.

Stages mirror the SIMDPolicy stage functions of the reference (internal/RendererCPU.h:39-189):
  GeneratePrimaryRays -> TraceRays -> ShadePrimary -> TraceShadowRays -> [TraceRays(lights) -> ShadeSecondary -> ...]
"""
import numpy as np
import pytest

from ray_b200 import capi, scenes
from common import Pair, bits_equal, by_xy, field_mismatch

pytestmark = pytest.mark.gpu

SCENES = {
    "cornell": lambda: scenes.cornell_box(96, 96),
    "zoo": lambda: scenes.material_zoo(),
    "zoo_env_dof": lambda: scenes.material_zoo(128, 96, lights=("rect", "sphere"), env=(0.4, 0.5, 0.7),
                                               filter=capi.FILTER_BLACKMAN_HARRIS, fstop=2.0),
    # hexagonal, rotated, anamorphic aperture + sensor shift + radiance clamps + exposure / gamma
    "zoo_lens_clamp": lambda: _lens_clamp_scene(),
    "instanced": lambda: scenes.instanced(36, 600, 128, 96),
    "hall_small": lambda: scenes.hall("principled", 160, 90, floor_res=48, n_columns=8, col_seg=12, col_rings=8,
                                      extra_lights=12),
    # every texture fetch on the path (base / roughness / metallic / specular / normal maps,
    # texture-driven Mix, alpha cut-out through the transparency loops, textured emissive triangles)
    "textured": lambda: scenes.textured(96, 72),
    # RGBE lat-long environment map: miss shading, quad-tree importance sampling + MIS,
    # a rotated background map for camera rays, a sky-portal rect light
    "envmap_zoo": lambda: scenes.envmap_zoo(96, 72),
    # the same textured scene built the way the reference does by default (settings_t::use_tex_compression = true):
    # BCn-coded blocks and YCoCg-coded base-colour maps (CoreRef.h:239-251, ShadeRef.cpp:1308,1411)
    "textured_compressed": lambda: scenes.textured(96, 72),
}


def _lens_clamp_scene():
    d = scenes.material_zoo(112, 80, lights=("spot", "disk", "line"), env=(0.2, 0.2, 0.25), filter=capi.FILTER_GAUSSIAN,
                            fstop=1.4)
    c = d.camera
    c.lens_blades, c.lens_rotation, c.lens_ratio = 6, 0.3, 1.3
    c.shift[0], c.shift[1] = 0.05, -0.03
    c.clamp_direct, c.clamp_indirect = 2.0, 1.0
    c.exposure, c.gamma = 0.5, 2.2
    c.filter_width = 2.0
    return d


@pytest.fixture(scope="module", params=list(SCENES))
def pair(request, oracle_mod):
    p = Pair(oracle_mod, SCENES[request.param](), tex_compression=(request.param == "textured_compressed"))
    p.name = request.param
    yield p
    p.close()


def _assert_records(a, b, what):
    a, b = by_xy(a), by_xy(b)
    assert len(a) == len(b), f"{what}: {len(a)} records vs reference {len(b)}"
    if not bits_equal(a, b):
        raise AssertionError(f"{what}: records differ bitwise in fields {field_mismatch(a, b)} of {len(a)}")


@pytest.mark.parametrize("iteration", [1, 7])
def test_generate_primary_rays(pair, iteration):
    p = pair.make_pass(iteration)
    rays, hits = pair.ctx.stage_generate_primary_rays(p)
    ref_rays, ref_hits = pair.osc.generate_primary_rays(pair.w, pair.h, (0, 0, pair.w, pair.h), iteration)
    assert len(rays) == pair.w * pair.h
    order, ref_order = np.argsort(rays["xy"], kind="stable"), np.argsort(ref_rays["xy"], kind="stable")
    assert bits_equal(rays[order], ref_rays[ref_order]), field_mismatch(rays[order], ref_rays[ref_order])
    assert bits_equal(hits[order], ref_hits[ref_order]), field_mismatch(hits[order], ref_hits[ref_order])


def test_trace_primary(pair):
    it = 3
    ref_rays, ref_hits = pair.osc.generate_primary_rays(pair.w, pair.h, (0, 0, pair.w, pair.h), it)
    o_rays, o_hits = pair.osc.trace_rays(it, ref_rays, ref_hits, False)
    g_rays, g_hits = pair.ctx.stage_trace_rays(pair.make_pass(it), ref_rays, ref_hits, False)
    # misses carry an unresolved prim_index by design but it is the same garbage on both sides
    assert bits_equal(g_hits, o_hits), field_mismatch(g_hits, o_hits)
    assert bits_equal(g_rays, o_rays), field_mismatch(g_rays, o_rays)
    assert (o_hits["v"] >= 0).any()


def test_wavefront_stage_by_stage(pair):
    """Walk 1 sample through all bounces, feeding BOTH sides the reference's outputs of the previous stage, and compare
    every stage's outputs bitwise: secondary rays, shadow rays, the radiance (temp) buffer and the primary AOVs."""
    it = 2
    w, h = pair.w, pair.h
    cam = pair.cam
    p = pair.make_pass(it)
    rays, hits = pair.osc.generate_primary_rays(w, h, (0, 0, w, h), it)
    rays, hits = pair.osc.trace_rays(it, rays, hits, False)

    temp = np.zeros((h, w, 4), np.float32)
    pair.ctx.fill_temp((0, 0, 0, 0))
    o_sec, o_sh, o_base, o_dn = pair.osc.shade(w, h, it, True, 0, rays, hits, temp)
    g_sec, g_sh = pair.ctx.stage_shade(p, True, 0, rays, hits)
    _assert_records(g_sec, o_sec, "primary shade: secondary rays")
    _assert_records(g_sh, o_sh, "primary shade: shadow rays")
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_TEMP), temp), "primary shade: colour buffer"
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_BASE_COLOR), o_base), "primary shade: base colour AOV"
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_DEPTH_NORMALS), o_dn), "primary shade: depth-normal AOV"

    pair.osc.trace_shadow_rays(w, it, o_sh, cam.clamp_direct, temp)
    pair.ctx.stage_trace_shadow_rays(p, o_sh, cam.clamp_direct)
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_TEMP), temp), "primary shadow: colour buffer"

    total = len(rays)
    sec = by_xy(o_sec)
    for bounce in range(1, cam.max_total_depth + 1):
        if len(sec) == 0:
            break
        hits0 = np.zeros(len(sec), dtype=hits.dtype)
        hits0["obj_index"] = -1
        hits0["prim_index"] = -1
        hits0["t"] = np.float32(3.402823466e+30)
        hits0["v"] = -1.0
        o_rays, o_hits = pair.osc.trace_rays(it, sec, hits0, True)
        g_rays, g_hits = pair.ctx.stage_trace_rays(p, sec, hits0, True)
        assert bits_equal(g_hits, o_hits), f"bounce {bounce} trace: hits {field_mismatch(g_hits, o_hits)}"
        assert bits_equal(g_rays, o_rays), f"bounce {bounce} trace: rays {field_mismatch(g_rays, o_rays)}"
        total += len(sec)

        o_sec, o_sh, _, _ = pair.osc.shade(w, h, it, False, bounce, o_rays, o_hits, temp)
        g_sec, g_sh = pair.ctx.stage_shade(p, False, bounce, o_rays, o_hits)
        _assert_records(g_sec, o_sec, f"bounce {bounce} shade: secondary rays")
        _assert_records(g_sh, o_sh, f"bounce {bounce} shade: shadow rays")
        assert bits_equal(pair.ctx.readback(capi.RC_BUF_TEMP), temp), f"bounce {bounce} shade: colour buffer"

        pair.osc.trace_shadow_rays(w, it, o_sh, cam.clamp_indirect, temp)
        pair.ctx.stage_trace_shadow_rays(p, o_sh, cam.clamp_indirect)
        assert bits_equal(pair.ctx.readback(capi.RC_BUF_TEMP), temp), f"bounce {bounce} shadow: colour buffer"
        sec = by_xy(o_sec)
    assert total > w * h, "no secondary rays were exercised"


@pytest.mark.parametrize("sort", [False, True])
def test_full_render_matches_reference_renderer(pair, oracle_mod, sort):
    """rc_render (the whole RenderScene sequence, with and without the results-neutral ray sort) against the
    reference's own Ref renderer run on the SAME wide-BVH scene object: north_star bar is 1e-4 per-pixel L-inf on the
    linear image; this backend is expected to be bit-identical."""
    spp = 4
    ref = oracle_mod.Renderer(capi.RT_REFERENCE, pair.w, pair.h)
    it = 0
    for _ in range(spp):
        it = ref.render(pair.osc, (0, 0, pair.w, pair.h), it)
    ref_raw, ref_final = ref.pixels(1), ref.pixels(0)
    ref_base, ref_dn = ref.pixels(2), ref.pixels(3)
    ref.close()

    pair.ctx.resize(pair.w, pair.h)
    pair.ctx.clear((0, 0, 0, 0))
    pair.ctx.fill_temp((0, 0, 0, 0))
    # fresh AOV accumulation: Resize() is a no-op at unchanged size, so run on a context-local clean state instead
    flags = 0 if sort else capi.RC_RENDER_NO_SORT
    for i in range(1, spp + 1):
        pair.ctx.render(pair.make_pass(i, flags=flags))
    raw = pair.ctx.readback(capi.RC_BUF_RAW)
    final = pair.ctx.readback(capi.RC_BUF_FINAL)
    diff = np.abs(raw - ref_raw)
    n_bad = int((diff.max(axis=-1) > 1e-4).sum())
    assert n_bad == 0, f"{n_bad} pixels differ by more than 1e-4 (L-inf {diff.max()})"
    assert bits_equal(raw, ref_raw), f"linear image not bit-identical: L-inf {diff.max()}, {int((diff > 0).any(-1).sum())} px"
    # the display transform goes through powf: the device runs a restatement of the host libm's algorithm (rt_math.cuh
    # libm_powf, tests/test_libm.py), so the tonemapped plane is bit-identical too
    assert bits_equal(final, ref_final), f"tonemapped image: L-inf {np.abs(final - ref_final).max()}"
    c = pair.ctx.counters()
    assert c["primary_rays"] >= spp * pair.w * pair.h
    if sort:
        # AOVs (running means of base colour and depth / normals, ShadeRef.cpp:1677-1698); the stage test above left
        # them dirty, so compare a run that starts from zeroed planes
        pair.ctx.resize(pair.w + 1, pair.h)
        pair.ctx.resize(pair.w, pair.h)
        for i in range(1, spp + 1):
            pair.ctx.render(pair.make_pass(i, flags=flags))
        assert bits_equal(pair.ctx.readback(capi.RC_BUF_BASE_COLOR), ref_base), "base colour AOV"
        assert bits_equal(pair.ctx.readback(capi.RC_BUF_DEPTH_NORMALS), ref_dn), "depth-normals AOV"


def test_adaptive_sampling_matches_reference_renderer(oracle_mod):
    """variance estimate + required_samples (RendererCPU.h:607-658): pixels whose two half-buffers agree stop being
    sampled after min_samples, raygen skips them (CoreRef.cpp:1446-1449).  Same image, bit for bit, as RendererRef."""
    desc = scenes.cornell_box(64, 64)
    desc.camera.min_samples = 4
    desc.camera.variance_threshold = 0.02
    pair = Pair(oracle_mod, desc)
    spp = 12
    ref = oracle_mod.Renderer(capi.RT_REFERENCE, pair.w, pair.h)
    it = 0
    for _ in range(spp):
        it = ref.render(pair.osc, (0, 0, pair.w, pair.h), it)
    ref_raw = ref.pixels(1)
    ref.close()
    pair.ctx.clear((0, 0, 0, 0))
    for i in range(1, spp + 1):
        pair.ctx.render(pair.make_pass(i))
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_RAW), ref_raw)
    c = pair.ctx.counters()
    assert c["primary_rays"] < spp * pair.w * pair.h, "no pixel converged: the adaptive path was not exercised"
    pair.close()


def test_nlm_denoise_matches_reference_renderer(oracle_mod):
    """RendererBase::DenoiseImage(region) (NLM half): same 8 spp accumulated on both sides (bit-identical,
    see above), then the joint NLM filter.  The filtered LINEAR image must be bit-identical (the weights go through a
    restated libm expf); the tonemapped plane goes through powf (tolerance as for rc_render).  A sub-rect call checks
    the clamped fetches at region borders that are not image borders."""
    desc = scenes.cornell_box(96, 80)
    pair = Pair(oracle_mod, desc)
    spp = 8
    ref = oracle_mod.Renderer(capi.RT_REFERENCE, pair.w, pair.h)
    it = 0
    for _ in range(spp):
        it = ref.render(pair.osc, (0, 0, pair.w, pair.h), it)
    pair.ctx.clear((0, 0, 0, 0))
    for i in range(1, spp + 1):
        pair.ctx.render(pair.make_pass(i))
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_RAW), ref.pixels(1))
    for rect in ((0, 0, pair.w, pair.h), (17, 9, 40, 33)):
        ref.denoise(rect, it)
        pair.ctx.denoise_nlm(rect, it)
        ref_raw, ref_final = ref.pixels(1), ref.pixels(0)
        raw, final = pair.ctx.readback(capi.RC_BUF_RAW), pair.ctx.readback(capi.RC_BUF_FINAL)
        x, y, w, h = rect
        sl = (slice(y, y + h), slice(x, x + w))
        assert np.isfinite(raw[sl]).all()
        d = np.abs(raw[sl] - ref_raw[sl])
        assert bits_equal(raw[sl], ref_raw[sl]), f"rect {rect}: filtered linear image L-inf {d.max()}, {int((d > 0).any(-1).sum())} px"
        assert bits_equal(final[sl], ref_final[sl]), f"rect {rect}: tonemapped L-inf {np.abs(final[sl] - ref_final[sl]).max()}"
    ref.close()
    pair.close()


@pytest.mark.parametrize("view_transform", [1, 2, 6, 9])  # AgX, AgX_Punchy, Filmic_MediumContrast, Filmic_VeryHighContrast
def test_lut_view_transforms_match_reference_renderer(oracle_mod, view_transform):
    """camera_desc_t::view_transform = AgX / Filmic (TonemapFilmic, TonemapRef.cpp:29-66): the 48^3 table comes from the
    reference through rc_set_view_lut; the tonemapped plane after rc_render and after the NLM denoiser is bit-identical,
    with a non-unit gamma on top.  Without the table the render call fails loudly."""
    desc = scenes.cornell_box(80, 64)
    desc.camera.view_transform = view_transform
    desc.camera.gamma = 1.8
    desc.camera.exposure = 0.5
    pair = Pair(oracle_mod, desc)
    assert pair.cam.view_transform == view_transform
    with pytest.raises(Exception):
        pair.ctx.render(pair.make_pass(1))
    pair.ctx.set_view_lut(view_transform, oracle_mod.view_lut(view_transform))
    spp = 6
    ref = oracle_mod.Renderer(capi.RT_REFERENCE, pair.w, pair.h)
    it = 0
    for _ in range(spp):
        it = ref.render(pair.osc, (0, 0, pair.w, pair.h), it)
    pair.ctx.clear((0, 0, 0, 0))
    for i in range(1, spp + 1):
        pair.ctx.render(pair.make_pass(i))
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_RAW), ref.pixels(1))
    final, ref_final = pair.ctx.readback(capi.RC_BUF_FINAL), ref.pixels(0)
    assert bits_equal(final, ref_final), f"tonemapped plane: L-inf {np.abs(final - ref_final).max()}"
    assert final[..., :3].std() > 0.01
    rect = (0, 0, pair.w, pair.h)
    ref.denoise(rect, it)
    pair.ctx.denoise_nlm(rect, it)
    assert bits_equal(pair.ctx.readback(capi.RC_BUF_FINAL), ref.pixels(0))
    ref.close()
    pair.close()


def test_unet_denoise_matches_reference_renderer(oracle_mod):
    """RendererBase::DenoiseImage(pass, region) (UNet half): same 8 spp accumulated on both sides
    (bit-identical), then the 16-pass UNet with the reference's own weight set handed over through rc_unet_set_weights.
    The fp32 path sums the same products in another order than the reference's 4-lane partial sums (and uses FMA), so the
    filtered linear image agrees to rounding noise, not bitwise: tolerance 2e-4 relative to (1 + |value|)."""
    desc = scenes.cornell_box(112, 80)  # 112 = 7 x 16, 80 = 5 x 16; a second case below is not a multiple of 16
    for (w, h) in ((112, 80), (100, 70)):
        desc = scenes.cornell_box(w, h)
        pair = Pair(oracle_mod, desc)
        spp = 8
        ref = oracle_mod.Renderer(capi.RT_REFERENCE, w, h)
        it = 0
        for _ in range(spp):
            it = ref.render(pair.osc, (0, 0, w, h), it)
        pair.ctx.clear((0, 0, 0, 0))
        for i in range(1, spp + 1):
            pair.ctx.render(pair.make_pass(i))
        assert bits_equal(pair.ctx.readback(capi.RC_BUF_RAW), ref.pixels(1))
        ref.denoise_unet((0, 0, w, h), it)
        pair.ctx.unet_set_weights(oracle_mod.unet_layers())
        pair.ctx.denoise_unet((0, 0, w, h), flags=capi.RC_UNET_FP32)
        ref_raw, ref_final = ref.pixels(1), ref.pixels(0)
        raw, final = pair.ctx.readback(capi.RC_BUF_RAW), pair.ctx.readback(capi.RC_BUF_FINAL)
        assert np.isfinite(raw).all()
        err = np.abs(raw[..., :3] - ref_raw[..., :3]) / (1.0 + np.abs(ref_raw[..., :3]))
        assert err.max() <= 2e-4, f"{w}x{h}: UNet (fp32) filtered image differs: max rel {err.max():g}"
        assert np.abs(final[..., :3] - ref_final[..., :3]).max() <= 1e-3
        # the filter actually filtered: the output is not the noisy input
        assert np.abs(raw[..., :3] - pair.ctx.readback(capi.RC_BUF_FULL)[..., :3]).mean() > 1e-4
        ref.close()
        pair.close()


def test_unet_tensor_core_path_matches_reference_renderer(oracle_mod):
    """The same UNet through the tcgen05 path (fp16 operands, fp32 accumulation in TMEM, rt_unet_tc.cuh) against the
    reference's fp32 CPU filter.  Activations are rounded to fp16 between the 16 layers (as on the reference's own GPU
    path), so the bar is fp16-level agreement: max relative error 3e-2 of (1 + |value|), mean 2e-3, and > 40 dB PSNR
    against the fp32 device path on the tonemapped image."""
    for (w, h) in ((160, 96), (100, 70)):
        desc = scenes.cornell_box(w, h)
        pair = Pair(oracle_mod, desc)
        spp = 8
        ref = oracle_mod.Renderer(capi.RT_REFERENCE, w, h)
        it = 0
        for _ in range(spp):
            it = ref.render(pair.osc, (0, 0, w, h), it)
        pair.ctx.clear((0, 0, 0, 0))
        for i in range(1, spp + 1):
            pair.ctx.render(pair.make_pass(i))
        ref.denoise_unet((0, 0, w, h), it)
        ref_raw = ref.pixels(1)
        pair.ctx.unet_set_weights(oracle_mod.unet_layers())
        pair.ctx.denoise_unet((0, 0, w, h), flags=capi.RC_UNET_FP32)
        f32_final = pair.ctx.readback(capi.RC_BUF_FINAL)
        pair.ctx.denoise_unet((0, 0, w, h), flags=capi.RC_UNET_TENSOR_CORES)
        raw, final = pair.ctx.readback(capi.RC_BUF_RAW), pair.ctx.readback(capi.RC_BUF_FINAL)
        assert np.isfinite(raw).all()
        err = np.abs(raw[..., :3] - ref_raw[..., :3]) / (1.0 + np.abs(ref_raw[..., :3]))
        mse = float(((final[..., :3] - f32_final[..., :3]) ** 2).mean())
        psnr = 10.0 * np.log10(1.0 / max(mse, 1e-12))
        print(f"unet tc {w}x{h}: max rel {err.max():.3g} mean rel {err.mean():.3g} PSNR vs fp32 path {psnr:.1f} dB")
        assert err.max() <= 3e-2 and err.mean() <= 2e-3, f"{w}x{h}: max rel {err.max():g}, mean rel {err.mean():g}"
        assert psnr > 40.0
        ref.close()
        pair.close()
