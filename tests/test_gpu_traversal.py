"""GPU: the two-phase traversal of the wavefront kernels (conservative classification per rectangle, the reference's own tests
only for candidates; ky_b200/csrc/kyd_device.cuh) answers every query exactly like the reference's list walk
(scene_t::intersect / occluded, ky.cpp:3172-3206).  The film tests check that through whole renders; this one asks the device
to compare both on rays built to sit on the classifier's decision boundaries."""
import pytest

import cases
import ky_b200 as ky

pytestmark = pytest.mark.gpu
TRAVERSAL = 2


@pytest.mark.parametrize("scene_key", ["cornell", "cornell_large_mirror_all_lights", "veach", "smallpt", "shapes", "ties_first", "ties_last",
                                       "inside_sphere_light", "inside_sphere_light_only"])
def test_two_phase_traversal_equals_the_list_walk(device, scene_key):
    if scene_key.startswith("ties_"):
        scene = cases.coplanar_tie_scene(scene_key[5:])
    elif scene_key.startswith("inside_sphere_light"):
        scene = cases.inside_sphere_light_scene(scene_key.endswith("_only"))   # ray origins far outside the rectangles' bounds
    else:
        scene = cases.make_scene(scene_key)
    device.upload(scene)
    for first in (0, 1 << 40):
        bad, hits = device.selftest(TRAVERSAL, first, 1 << 24)
        assert bad == 0, f"{bad} queries differ from the list walk"
        assert hits > (1 << 22)


def test_traversal_selftest_needs_a_scene():
    fresh = ky.Device(0)
    with pytest.raises(RuntimeError, match="needs an uploaded scene"):
        fresh.selftest(TRAVERSAL, 0, 16)
    fresh.close()
