"""Material-type dispatch of the ReSTIR statics (P/ReSTIRIntegrator.h:32-59) on a scene that holds EVERY MaterialType:

  getMaterialBRDFEvalFunc : PHONG, DIELECTRIC -> Phong lobe; LAMBERT and everything else (NORMAL, MIRROR,
                            DIELECTRIC_TRANSPARENT, UNSUPPORTED) -> MaterialLambert::evalBRDF
  getMaterialSampleFunc   : LAMBERT -> cosine; everything else -> MaterialPhong::sampleBRDF
  getMaterialPDFEvalFunc  : always MaterialPhong::evalPdf

Three tiers: (1) the oracle against the reference's OWN classes (MaterialMirror, MaterialNormal, MaterialTransparent,
MaterialDielectric compiled in place by oracle/ref_shim) bit for bit, (2) the product's kernel bodies (host emulation)
against the oracle, (3) -m gpu: the CUDA kernels through the C ABI against the oracle, inline and wavefront."""
import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
import ref_binding as rb
from restir_embree_b200 import Camera, abi, scenes
from test_emu_parity import ALL_BUFS, bits
from test_ref_pin import check_against

TYPES = [abi.MAT_NORMAL, abi.MAT_MIRROR, abi.MAT_DIELECTRIC, abi.MAT_DIELECTRIC_TRANSPARENT, abi.MAT_LAMBERT, abi.MAT_PHONG]


def all_types_scene():
    """the tiny room; its six wall materials and twelve blob materials cycle through the six material types"""
    sc = scenes.make_tiny_scene(seed=5, n_blobs=6)
    k = 0
    for m in sc.materials:
        if sum(m["emission"]) > 0:
            continue
        m["type"] = TYPES[k % len(TYPES)]
        m["ior"] = 1.5
        k += 1
    used = {sc.materials[s[2]]["type"] for s in sc.surfaces}
    assert used >= set(TYPES)
    return sc


PARAMS = [
    dict(M_Area=4, M_Brdf=2, doSpatialReuse=1, doTemporalReuse=1),
    dict(M_Area=3, M_Brdf=1, doSpatialReuse=1, doTemporalReuse=1, doVisibilityPass=1, spatialWeightCalc=2),
]
W, H = 48, 32


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs the reference checkout at build time)")
@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_oracle_matches_reference_classes_for_every_material_type(pi):
    sc = all_types_scene()
    ref = rb.Reference(W, H, sc)
    p = abi.default_params(**PARAMS[pi])
    ref.set_params(p)
    o = ob.Oracle(W, H, seed=123, rng=ob.RNG_LEGACY, math=ob.MATH_LIBM, tracer=ob.TRACER_BRUTE, cache_iim=0)
    o.upload_scene(sc)
    o.set_params(p)
    seen = set()
    for f in range(3):
        cam = ref.camera(60.0, (2.2 + 0.05 * f, -2.4, 1.4), (0, 0, 1.0))
        a = ref.produce_restir()
        b = o.render_frame(cam, f)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        g = ref.gbuffer()
        check_against(o, a, ref.reservoirs(), g)
        seen |= set(np.unique(g[..., 17].astype(np.int32)).tolist())
    assert seen >= set(TYPES), f"material types in the reference's G-buffer: {sorted(seen)}"


WAVE = [0, 1]


@pytest.mark.parametrize("wave", WAVE)
def test_emulated_kernels_match_oracle_for_every_material_type(wave):
    sc = all_types_scene()
    p = abi.default_params(**PARAMS[wave], wavefront=wave, lightSampler=wave)
    o = ob.Oracle(W, H, seed=11, tracer=ob.TRACER_BRUTE)
    e = eb.Emu(W, H, seed=11)
    for x in (o, e):
        x.upload_scene(sc)
        x.set_params(p)
    for f in range(3):
        cam = Camera(W, H, 60, (2.2 + 0.05 * f, -2.4, 1.4), (0, 0, 1.0))
        a, b = o.render_frame(cam, f), e.render_frame(cam, f)
        assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
        for buf in ALL_BUFS:
            assert np.array_equal(bits(o.readback(buf)), bits(e.readback(buf))), (f, buf)
    types = e.readback(abi.BUF_GBUF_SPEC_TYPE)[..., 3].view(np.uint32) & 0xFF
    assert set(np.unique(types).tolist()) >= set(TYPES)


@pytest.mark.gpu
@pytest.mark.parametrize("wave", WAVE)
def test_gpu_matches_oracle_for_every_material_type(gpu, wave):
    from restir_embree_b200.renderer import Renderer
    sc = all_types_scene()
    p = abi.default_params(**PARAMS[wave], wavefront=wave, lightSampler=wave)
    o = ob.Oracle(W, H, seed=11, tracer=ob.TRACER_BRUTE)
    o.upload_scene(sc)
    o.set_params(p)
    with Renderer(W, H, seed=11) as r:
        r.upload_scene(sc)
        r.set_params(p)
        for f in range(3):
            cam = Camera(W, H, 60, (2.2 + 0.05 * f, -2.4, 1.4), (0, 0, 1.0))
            a, b = o.render_frame(cam, f), r.render_frame(cam, f)
            assert np.array_equal(bits(a), bits(b)), f"frame {f}: {(a != b).any(-1).sum()} px differ"
            for buf in ALL_BUFS:
                assert np.array_equal(bits(o.readback(buf)), bits(r.readback(buf))), (f, buf)
