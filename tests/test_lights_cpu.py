"""CPU-only: light creation (rad/lightmap/lights.go:38-426, common/types/entity.go:96-158, vmath/quadratic/solver.go) --
the producer of K3's `vrad_light` records.  The product's host functions against the oracle's function-by-function
restatement (byte for byte) and against hand-derived values."""
import math

import numpy as np
import pytest

from oracle import pyoracle
from vrad_b200 import scenes
from vrad_b200.environment import VradError, light_for_string, lights_from_entities, lights_from_patches
from vrad_b200.lib import LIGHT_ENTITY_DTYPE

POINT, SPOT, ENV = 0, 1, 2


def _ent(classname, origin=(0, 0, 0), light="255 255 255 200", **kw):
    e = np.zeros(1, LIGHT_ENTITY_DTYPE)[0]
    e["classname"] = classname; e["origin"] = origin
    try:
        e["light"] = light_for_string(light); e["light_ok"] = 1
    except VradError:
        e["light_ok"] = 0
    for k, v in kw.items():
        e[k] = v
    return e


def test_light_for_string():
    assert np.allclose(light_for_string("255 255 255 200"), 200.0)                     # white, brightness 200
    assert np.allclose(light_for_string("255"), 255.0) and np.allclose(light_for_string("255 255 255"), 255.0)
    v = light_for_string("128 64 255 100")                                             # gamma 2.2 to linear, then * 100/255
    assert np.allclose(v, [((128 / 255) ** 2.2) * 255 * 100 / 255, ((64 / 255) ** 2.2) * 255 * 100 / 255, 100.0], rtol=1e-6)
    assert np.allclose(light_for_string("255 255 255 200 255 255 255 800"), 200.0)     # LDR tuple of an LDR+HDR pair
    for bad in ("", "1 2", "-1 0 0 10", "abc"):
        with pytest.raises(VradError):
            light_for_string(bad)
        assert pyoracle.light_for_string(bad)[1] != 0
    for ok in ("255 255 255 200", "17", "10 20 30", "128 64 255 100", "1 2 3 4 5 6 7 8"):
        assert light_for_string(ok).tobytes() == pyoracle.light_for_string(ok)[0].tobytes()


def test_point_light_legacy_attenuation():
    # no attenuation keys at all -> constant 1, intensity unchanged (ratio 1)
    L = lights_from_entities([_ent(POINT, (10, 20, 30))])[0]
    assert L["type"] == scenes.EMIT_POINT and tuple(L["origin"]) == (10, 20, 30)
    assert (L["constant_attn"], L["linear_attn"], L["quadratic_attn"]) == (1.0, 0.0, 0.0) and np.allclose(L["intensity"], 200.0)
    assert (L["start_fade"], L["end_fade"], L["cap_dist"]) == (0.0, -1.0, np.float32(1e22))
    # quadratic only: "scale intensity for unit 100 distance" (lights.go:335-339) -> x 10000
    L = lights_from_entities([_ent(POINT, quadratic_attn=1.0)])[0]
    assert np.allclose(L["intensity"], 200.0 * 10000.0) and L["quadratic_attn"] == 1.0 and L["constant_attn"] == 0.0
    # values below EQUAL_EPSILON are clamped to 0
    L = lights_from_entities([_ent(POINT, constant_attn=0.0005, linear_attn=0.5)])[0]
    assert L["constant_attn"] == 0.0 and L["linear_attn"] == 0.5 and np.allclose(L["intensity"], 200.0 * 50.0)


def test_spot_light_cones_and_normal():
    L = lights_from_entities([_ent(SPOT, inner_cone=30.0, cone=45.0, exponent=2.0, pitch=-90.0)])[0]
    assert L["type"] == scenes.EMIT_SPOTLIGHT and L["exponent"] == 2.0
    assert L["stopdot"] == np.float32(math.cos(math.radians(30))) and L["stopdot2"] == np.float32(math.cos(math.radians(45)))
    assert np.allclose(L["normal"], [0, 0, -1], atol=1e-6)                              # pitch -90: straight down
    # defaults: inner cone 10 degrees, outer = inner; outer never smaller than inner
    L = lights_from_entities([_ent(SPOT), _ent(SPOT, inner_cone=40.0, cone=20.0)])
    assert L[0]["stopdot"] == L[0]["stopdot2"] == np.float32(math.cos(math.radians(10)))
    assert L[1]["stopdot"] == L[1]["stopdot2"] == np.float32(math.cos(math.radians(40)))
    # 180/180 is a point light; angles above 90 are clamped
    L = lights_from_entities([_ent(SPOT, inner_cone=180.0, cone=180.0, exponent=3.0), _ent(SPOT, inner_cone=100.0, cone=120.0)])
    assert L[0]["type"] == scenes.EMIT_POINT and L[0]["stopdot"] == 0 and L[0]["exponent"] == 0
    assert abs(L[1]["stopdot"]) < 1e-6 and abs(L[1]["stopdot2"]) < 1e-6
    # target entity wins over angles; yaw/pitch from "angles" when angle/pitch are 0; ANGLE_UP
    L = lights_from_entities([_ent(SPOT, (0, 0, 100), has_target=1, target_origin=(30, 40, 100)),
                              _ent(SPOT, angles=(-30.0, 90.0, 0.0)), _ent(SPOT, angle=-1.0, pitch=90.0)])
    assert np.allclose(L[0]["normal"], [0.6, 0.8, 0.0], atol=1e-6)
    assert np.allclose(L[1]["normal"], [0.0, math.cos(math.radians(30)), -0.5], atol=1e-6)
    assert np.allclose(L[2]["normal"], [0.0, 0.0, 1.0], atol=1e-6)


def test_fifty_percent_distance_falloff():
    d50, d0 = 200.0, 800.0
    L = lights_from_entities([_ent(POINT, fifty_percent_distance=d50, zero_percent_distance=d0)])[0]
    a, b, c = float(L["quadratic_attn"]), float(L["linear_attn"]), float(L["constant_attn"])
    f = lambda d: 1.0 / (c + d * (b + d * a))
    assert abs(f(d50) - 0.5) < 1e-5                                                     # half brightness at d50 (after the rescale)
    assert 2 * a + b >= 0 and f(d0) < f(d50) / 10 and f(0.0) > f(d50)                   # the fit is kept monotonic at x = 1 (solver.go:27-41)
    assert np.allclose(L["intensity"], 200.0)                                           # no unit-100 rescale on this branch
    if a != 0 and b / (-2 * a) > 0:
        assert L["cap_dist"] == np.float32(b / (-2 * a)) and L["end_fade"] == np.float32(10 * L["cap_dist"])
    else:
        assert L["cap_dist"] == np.float32(1e22) and L["end_fade"] == -1.0
    # hard falloff: fade from 3/4 of the way to the zero distance
    H = lights_from_entities([_ent(POINT, fifty_percent_distance=d50, zero_percent_distance=d0, hardfalloff=1)])[0]
    assert H["end_fade"] == d0 and H["start_fade"] == np.float32(0.75 * d0 + 0.25 * d50)
    # zero distance below the fifty distance is replaced by 2 * d50
    Z = lights_from_entities([_ent(POINT, fifty_percent_distance=d50, zero_percent_distance=10.0, hardfalloff=1)])[0]
    assert Z["end_fade"] == 2 * d50


def test_light_environment_and_order():
    ents = [_ent(POINT, (1, 1, 1)), _ent(ENV, (0, 0, 500), "255 255 255 100", pitch=-60.0, angles=(0.0, 45.0, 0.0)),
            _ent(ENV, (9, 9, 9), "10 10 10 10"), _ent(SPOT, (2, 2, 2))]
    L = lights_from_entities(ents)
    assert [int(t) for t in L["type"]] == [scenes.EMIT_POINT, scenes.EMIT_SKYLIGHT, scenes.EMIT_SKYAMBIENT, scenes.EMIT_SPOTLIGHT]   # 2nd environment ignored
    assert np.allclose(L[1]["intensity"], 100.0) and np.allclose(L[2]["intensity"], 50.0)      # no "_ambient": half of the sun
    assert np.allclose(L[1]["normal"], [math.cos(math.radians(45)) * 0.5, math.sin(math.radians(45)) * 0.5, -math.sin(math.radians(60))], atol=1e-6)
    amb = _ent(ENV, (0, 0, 500), "255 255 255 100", ambient_ok=1, ambient=(7.0, 8.0, 9.0))
    assert tuple(lights_from_entities([amb])[1]["intensity"]) == (7.0, 8.0, 9.0)
    with pytest.raises(VradError):
        lights_from_entities([_ent(7)])                                                  # unknown class


def test_entities_match_oracle_bytewise():
    rng = scenes.SplitMix64(2024)
    n = 400
    ents = np.zeros(n, LIGHT_ENTITY_DTYPE)
    ents["classname"] = rng.integers(n, 3)
    ents["origin"] = rng.uniform(3 * n, -2000, 2000).reshape(n, 3)
    ents["light"] = rng.uniform(3 * n, 0, 400).reshape(n, 3); ents["light_ok"] = rng.integers(n, 5) > 0
    ents["has_target"] = rng.integers(n, 3) == 0
    ents["target_origin"] = rng.uniform(3 * n, -2000, 2000).reshape(n, 3)
    ents["angles"] = rng.uniform(3 * n, -180, 180).reshape(n, 3)
    ents["pitch"] = np.where(rng.integers(n, 2) == 0, 0, rng.uniform(n, -90, 90))
    ents["angle"] = np.choose(rng.integers(n, 4), [np.zeros(n), np.full(n, -1.0), np.full(n, -2.0), rng.uniform(n, 0, 360)])
    ents["inner_cone"] = np.choose(rng.integers(n, 4), [np.zeros(n), rng.uniform(n, 1, 120), np.full(n, 180.0), rng.uniform(n, 1, 60)])
    ents["cone"] = np.choose(rng.integers(n, 4), [np.zeros(n), rng.uniform(n, 1, 120), np.full(n, 180.0), rng.uniform(n, 1, 60)])
    ents["exponent"] = rng.uniform(n, 0, 3)
    ents["fifty_percent_distance"] = np.where(rng.integers(n, 2) == 0, 0, rng.uniform(n, 10, 1000))
    ents["zero_percent_distance"] = rng.uniform(n, 0, 4000)
    ents["hardfalloff"] = rng.integers(n, 2)
    for k in ("constant_attn", "linear_attn", "quadratic_attn"):
        ents[k] = np.where(rng.integers(n, 2) == 0, 0, rng.uniform(n, 0, 2))
    ents["distance"] = rng.uniform(n, 0, 500)
    ents["ambient_ok"] = rng.integers(n, 2); ents["ambient"] = rng.uniform(3 * n, 0, 100).reshape(n, 3)
    got = lights_from_entities(ents)
    want = pyoracle.lights_from_entities(ents, scenes.LIGHT_DTYPE)
    assert got.shape == want.shape and got.tobytes() == want.tobytes()
    assert np.isfinite(got["intensity"]).all() and (got["type"] == scenes.EMIT_SKYLIGHT).sum() == 1


def test_surface_lights_from_patches():
    n = 50
    rng = scenes.SplitMix64(9)
    origin = rng.uniform(3 * n, -100, 100).reshape(n, 3); normal = np.tile(np.float32([0, 0, 1]), (n, 1))
    base = np.zeros((n, 3), np.float32); base[::5] = rng.uniform(3 * 10, 0.5, 3.0).reshape(10, 3); base[1] = 0.05   # below the threshold
    area = rng.uniform(n, 100, 1000); scale2 = rng.uniform(2 * n, 0.2, 0.3).reshape(n, 2); base_area = rng.uniform(n, 50, 100)
    base_area[5] = 0.0                                                                   # no base area: skipped
    child1 = np.full(n, -1, np.int32); child1[10] = 11                                   # a parent patch: skipped
    got = lights_from_patches(origin, normal, base, area, scale2, base_area, child1)
    want = pyoracle.lights_from_patches(origin, normal, base, area, scale2, base_area, child1, 0.1, scenes.LIGHT_DTYPE)
    assert got.tobytes() == want.tobytes() and len(got) == 8 and np.all(got["type"] == scenes.EMIT_SURFACE)
    k = 15
    s = np.float32(1.0) * area[k] * scale2[k, 0] * scale2[k, 1] / base_area[k]
    idx = [i for i in range(0, n, 5) if i not in (5, 10)].index(k)
    assert np.array_equal(got[idx]["intensity"], (base[k] * s) * np.float32(10000.0))
    assert tuple(got[idx]["origin"]) == tuple(origin[k]) and got[idx]["end_fade"] == -1.0
