import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # no test may hang the run: with pytest-timeout installed, 15 minutes per test unless the command line says otherwise
    if config.pluginmanager.hasplugin("timeout") and not getattr(config.option, "timeout", None):
        config.option.timeout = 900


@pytest.fixture(scope="session")
def s1_scene():
    from vrad_b200 import scenes
    return scenes.box_room()


@pytest.fixture(scope="session")
def s1_oracle(s1_scene):
    from oracle import pyoracle
    return pyoracle.env_from_scene(s1_scene)


@pytest.fixture(scope="session")
def s2_small_scene():
    """3x2-room cut of the S2 multi-room map (same generator, ~2.3k triangles, ~8.6k patches)."""
    from vrad_b200 import scenes
    return scenes.multi_room(nx=3, ny=2, boxes_per_room=30)


@pytest.fixture(scope="session")
def s2_small_oracle(s2_small_scene):
    from oracle import pyoracle
    return pyoracle.env_from_scene(s2_small_scene)


@pytest.fixture(scope="session")
def s1_gpu(s1_scene):
    from vrad_b200.environment import environment_from_scene
    env = environment_from_scene(s1_scene)
    yield env
    env.close()


@pytest.fixture(scope="session")
def s2_small_gpu(s2_small_scene):
    from vrad_b200.environment import environment_from_scene
    env = environment_from_scene(s2_small_scene)
    yield env
    env.close()
