import glob
import os

import numpy as np

import scenes

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_SCENES = {
    "default_160x90_d32": scenes.default_scene,
    "default_moved_192x108_d8": scenes.default_scene,
    "default_above_128x128_d3_spp4": scenes.default_scene,
    "small12_seed1_160x100_d8": lambda: scenes.small_random_scene(12, 1),
    "small40_seed3_160x100_d8": lambda: scenes.small_random_scene(40, 3),
}


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def channels(px):
    px = np.asarray(px).astype(np.int64)
    return np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], axis=-1)


def assert_image_parity(got, ref, what=""):
    """BASELINE.json tolerance: per-channel error <= 1/255 on >= 99.9 % of pixels, no pixel off by more than 4/255."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    d = np.abs(channels(got) - channels(ref)).max(-1)
    frac_gt1 = float((d > 1).mean())
    assert d.max() <= 4, "%s: max channel error %d/255" % (what, d.max())
    assert frac_gt1 <= 1e-3, "%s: %.4f%% of pixels differ by more than 1/255" % (what, 100 * frac_gt1)
    return int((got != ref).sum())
