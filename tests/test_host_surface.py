"""CPU: the C++ host class surface (include/ky.hpp, include/ky_entry.hpp): flattening and the shapes
of the reference-named entry points."""
import numpy as np
import pytest

import cases
import ky_b200 as ky


def test_cornell_default_scene_layout():
    # 12 surfaces, 1 light, surface order left,right,top,bottom(glossy),back,mirror ball,glass ball,4 white sides,emitter
    s = cases.make_scene("cornell")
    d = s.desc
    assert (d.surface_count, d.light_count, d.shape_count, d.material_count, d.environment_light) == (12, 1, 13, 8, -1)
    kinds = [s.materials[sf.material].kind for sf in s.surfaces]
    assert kinds == [ky.MAT_MATTE, ky.MAT_MATTE, ky.MAT_MATTE, ky.MAT_PLASTIC, ky.MAT_MATTE, ky.MAT_MIRROR, ky.MAT_GLASS] + [ky.MAT_MATTE] * 5
    assert [sf.area_light for sf in s.surfaces] == [-1] * 11 + [0]
    assert s.lights[0].kind == ky.LIGHT_AREA and s.lights[0].shape == s.surfaces[11].shape
    assert list(s.lights[0].color) == [25, 25, 25]
    glossy = s.materials[s.surfaces[3].material]
    assert glossy.exponent == 90 and abs(glossy.specular_probability - 0.875) < 1e-6


def test_veach_scene_keeps_the_reference_cross_wiring():
    # ky.cpp:3498-3499 vs 3525-3526: light1 samples ball2's shape but sits on ball1's surface, and vice versa
    s = cases.make_scene("veach")
    d = s.desc
    assert (d.surface_count, d.light_count) == (11, 5)
    surf_ball1, surf_ball2 = s.surfaces[7], s.surfaces[8]
    assert surf_ball1.area_light == 1 and surf_ball2.area_light == 2
    assert s.lights[1].shape == surf_ball2.shape and s.lights[2].shape == surf_ball1.shape
    assert all(s.materials[s.surfaces[i].material].exponent == 5000 for i in range(2, 6))
    assert [round(l.color[0], 3) for l in s.lights] == [800.0, 901.803, 100.0, 11.111, 1.235]


def test_environment_and_direction_lights_carry_the_preprocessed_world_radius():
    s = cases.make_scene("cornell_large_mirror_all_lights")
    kinds = [l.kind for l in s.lights]
    assert kinds == [ky.LIGHT_AREA, ky.LIGHT_DIRECTION, ky.LIGHT_POINT, ky.LIGHT_ENVIRONMENT]
    assert s.desc.environment_light == 3
    r = s.lights[1].world_radius
    assert r > 1 and r == s.lights[3].world_radius
    n = np.array(list(s.lights[1].direction))
    assert abs(np.linalg.norm(n) - 1) < 1e-6


def test_both_large_balls_is_an_error():
    with pytest.raises(RuntimeError, match="both large balls"):
        ky.Scene(ky.SCENE_CORNELL, 8, 8, ky.CB_LARGE_MIRROR | ky.CB_LARGE_GLASS)


@pytest.mark.parametrize("name,shape", [
    ("render_single_scene", (1024, 1024)), ("render_debug", (308, 1536)), ("render_multiple_integrator", (1024, 1280)),
    ("render_direct_sample_enum", (1024, 1280)), ("render_multiple_scene", (768, 1024)), ("render_mis_scene", (616, 1536)),
    ("render_lighting_enum", (256, 1024))])
def test_entry_point_film_shapes_are_the_reference_ones(name, shape):
    # README figure sizes match the film_grid_t shapes
    film = ky.render_entry(name, render=False)
    assert film.shape[:2] == shape
