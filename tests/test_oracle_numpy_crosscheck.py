"""The C++ oracle against a second, independently written restatement of RayTracer.cs (tests/numpy_ref.py: vectorised numpy
fp32, literal two-root sphere test, no recursion) at ReflectionRecursionLimit = 0: pixels, primary hit ids and the bit
patterns of the primary distances must all agree.  Neither is the C# program (it cannot run here), but two translations
written in different styles agreeing bit for bit rules out transcription slips in operation order."""
import warnings

import numpy as np
import pytest

import numpy_ref
import oracle_lib as O
import scenes


@pytest.mark.parametrize("name,camkw", [
    ("default", dict()), ("default", dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15)), ("default", dict(pos=(0, 4.0, 6.0), pitch=1.2)),
    ("small12", dict(pos=(0, 1.5, -4.0), pitch=0.1)), ("small40", dict(pos=(1.0, 2.5, -3.0), yaw=-0.2, pitch=0.3))])
def test_oracle_equals_numpy_restatement(built, name, camkw):
    sc = {"default": scenes.default_scene, "small12": lambda: scenes.small_random_scene(12, 1),
          "small40": lambda: scenes.small_random_scene(40, 3)}[name]()
    w, h = 256, 144
    cam = scenes.make_camera(width=w, height=h, **camkw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        px, code, t = numpy_ref.render_cap0(sc, cam, w, h)
    r = O.render(sc, cam, w, h, 0, want_aov=True)
    assert np.array_equal(code, r["aov_id"])
    assert np.array_equal(t.view(np.uint32), r["aov_t"].view(np.uint32))
    assert np.array_equal(px, r["pixels"]), "%d pixels differ" % (px != r["pixels"]).sum()
    assert (px != 0).mean() > 0.3


@pytest.mark.parametrize("name,camkw,cap", [
    ("default", dict(), 1), ("default", dict(), 8), ("default", dict(pos=(0.3, 0.5, -1.0), yaw=0.2, pitch=0.15), 3),
    ("default", dict(pos=(-2.0, 2.5, 3.0), yaw=-0.4, pitch=0.5), 32), ("small12", dict(pos=(0, 1.5, -4.0), pitch=0.1), 4),
    ("small40", dict(pos=(1.0, 2.5, -3.0), yaw=-0.2, pitch=0.3), 2)])
def test_oracle_equals_numpy_restatement_with_recursion(built, name, camkw, cap):
    """The same at recursion limits > 0: numpy_ref.render follows TraceSecondaryRay on index subsets (mirror chains, the
    too-close / beyond-the-limit terminal colours, shading of reflected hits)."""
    sc = {"default": scenes.default_scene, "small12": lambda: scenes.small_random_scene(12, 1),
          "small40": lambda: scenes.small_random_scene(40, 3)}[name]()
    w, h = 192, 108
    cam = scenes.make_camera(width=w, height=h, **camkw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        px, code, t = numpy_ref.render(sc, cam, w, h, cap)
    r = O.render(sc, cam, w, h, cap, want_aov=True)
    assert np.array_equal(code, r["aov_id"])
    assert np.array_equal(t.view(np.uint32), r["aov_t"].view(np.uint32))
    assert np.array_equal(px, r["pixels"]), "%d pixels differ" % (px != r["pixels"]).sum()
    if cap >= 1:
        r0 = O.render(sc, cam, w, h, 0)
        assert (r0["pixels"] != r["pixels"]).any()          # the recursion does change the picture
