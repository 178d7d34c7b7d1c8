"""The random scenes of the fuzz parity tests (tests/test_gpu_fuzz.py), checked on the CPU: they load, they are
deterministic in their seed, they differ between seeds, and the oracle renders them with light reaching the film."""
import numpy as np
import pytest

import oracle_lib as O
from psychopath_b200 import api
from scenes import random_scene_text


def test_generator_is_deterministic_and_seeds_differ():
    assert random_scene_text(3) == random_scene_text(3)
    assert len({random_scene_text(s) for s in range(8)}) == 8


@pytest.mark.parametrize("seed", [0, 5, 9])
def test_random_scene_loads_and_the_oracle_renders_it(seed):
    scene = api.Scene.from_psy_string(random_scene_text(seed))
    d = scene.desc.contents
    assert (d.res_x, d.res_y, d.spp, d.seed) == (64, 48, 4, seed + 1)
    assert d.n_tris > 20 and d.n_light_instances >= 1
    img, st, _ = O.render_image(scene.desc, d.spp, d.seed, spb=4096)
    assert np.isfinite(img).all()
    assert (img[..., 1] > 0).mean() > 0.9
    assert st.ray_count > d.res_x * d.res_y * d.spp
