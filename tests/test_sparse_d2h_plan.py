"""Sparse device->host return (RT_OPT_SPARSE_D2H, plan_rows in csrc/rt_gate.cuh): rt_render does not copy what the frame gates prove
black and zero-fills it on the host instead. The plan is sound iff every pixel it leaves out is 0x00000000 in the oracle's frame
(RayTracer.cs:993: a primary ray that hits nothing). Checked here on the CPU for many cameras; the GPU tests then compare whole
frames with the option on and off."""
import numpy as np
import pytest

import hostemu_lib as E
import oracle_lib as O
import scenes


def _uncopied_mask(kind, rx0, rx1, w, h):
    m = np.zeros((h, w), bool)
    m[kind == E.ROW_BLACK, :] = True
    rect_rows = kind == E.ROW_RECT
    m[rect_rows, :rx0] = True
    m[rect_rows, rx1 + 1:] = True
    return m


def _check(sc, cam, w, h, depth=8):
    kind, rx0, rx1, sparse = E.row_plan(sc, cam, w, h)
    ref = O.render(sc, cam, w, h, depth)["pixels"]
    m = _uncopied_mask(kind, rx0, rx1, w, h)
    assert sparse == bool(m.any())
    assert not (ref[m] != 0).any(), "%d skipped pixels are not black" % int((ref[m] != 0).sum())
    return float(m.mean())


def test_default_camera_skips_a_third_of_the_frame(built):
    sc = scenes.default_scene()
    for w, h in ((1280, 720), (640, 360), (333, 187)):
        frac = _check(sc, scenes.make_camera(width=w, height=h), w, h)
        assert frac > 0.30, frac          # the sky above the horizon, beside and above the spheres


@pytest.mark.parametrize("camkw", [dict(pos=(0.0, 3.0, 2.0), pitch=1.3), dict(pos=(0.0, 0.5, 0.0), pitch=-0.6),
                                   dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), dict(pos=(-2.0, 2.5, 3.0), yaw=-0.4, pitch=0.5),
                                   dict(pos=(0.0, 0.0, 6.0), yaw=3.1), dict(pos=(0.0, -0.5, 0.0), pitch=-1.2),
                                   dict(pos=(0.0, -3.0, 0.0), pitch=0.3), dict(pos=(2.5, 0.0, 4.0)), dict(pos=(0.0, 40.0, 0.0), pitch=1.5)])
def test_cameras(built, camkw):
    """floor everywhere / mostly sky / tilted / behind the spheres / below the floor / inside a sphere / far above"""
    w, h = 320, 180
    _check(scenes.default_scene(), scenes.make_camera(width=w, height=h, **camkw), w, h)


@pytest.mark.parametrize("seed", range(10))
def test_random_scenes_and_cameras(built, seed):
    rng = np.random.default_rng(100 + seed)
    sc = scenes.default_scene() if seed % 3 == 0 else scenes.small_random_scene(int(rng.integers(0, 9)), seed)
    w, h = 208, 117
    for _ in range(5):
        pos = tuple(rng.uniform(-6, 6, 3) * np.array([1, 0.5, 1]) + np.array([0, 1.0, -3]))
        cam = scenes.make_camera(pos=pos, yaw=float(rng.uniform(-3.2, 3.2)), pitch=float(rng.uniform(-1.5, 1.5)), width=w, height=h)
        _check(sc, cam, w, h)


def test_tilted_plane_and_degenerate_scene_are_sound(built):
    """A plane whose horizon is not horizontal on screen (rows only partly sky are copied whole), and the degenerate scene."""
    sc = scenes.default_scene()
    sc.planes[0, 3:6] = np.float32([0.28, 0.96, 0.0])          # unit normal, tilted about z
    w, h = 240, 136
    for camkw in (dict(), dict(pos=(0.0, 1.0, -2.0), pitch=0.2), dict(pos=(1.0, 0.5, 0.0), yaw=1.0)):
        _check(sc, scenes.make_camera(width=w, height=h, **camkw), w, h)
    _check(scenes.degenerate_scene(), scenes.make_camera(width=160, height=96), 160, 96)


def test_no_plane_means_everything_outside_the_rectangle_is_skipped(built):
    sc = scenes.default_scene()
    sc = scenes.Scene(sc.spheres, np.zeros((0, 20), np.float32), sc.lights, sc.ambient)
    w, h = 320, 180
    frac = _check(sc, scenes.make_camera(width=w, height=h), w, h)
    assert frac > 0.6
