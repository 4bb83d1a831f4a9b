"""Scene and camera definitions for the ray-cast path: the reference's default scene and the synthetic
scaled scenes named in BASELINE.json `configs`.

All records are flat float32 arrays in the C# field order, which is also the C-ABI layout (include/rtb200.h):
  sphere : 18 floats  center[3], radius, Kd[3], Ka[3], Ks[3], n, Km[3], radius^2      (RayTracer.cs:308-338, :60-80)
  plane  : 20 floats  center[3], normal[3], Kd[3], Ka[3], Ks[3], n, Km[3], tiled      (RayTracer.cs:260-303)
  light  :  4 floats  position[3], intensity                                           (RayTracer.cs:236-255)

Pure numpy/python: no CUDA, no oracle imports. Deterministic (PCG32, stated seeds).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------------------
# Material factories — RayTracer.cs:117-158
# --------------------------------------------------------------------------------------------------------
def material(kd, ka, ks, n, km):
    return np.array([*kd, *ka, *ks, n, *km], dtype=F32)


def mat_diffuse(c):                       # :117-119
    return material(c, c, (0, 0, 0), 0.0, (0, 0, 0))


def mat_plastic(c, n=1.0):                # :127-129
    return material(c, c, (0.4, 0.4, 0.4), n, (0, 0, 0))


def mat_metal(c, n=1.0):                  # :137-139
    return material(c, c, c, n, (0, 0, 0))


def mat_mirror(c):                        # :146-148
    return material((0, 0, 0), (0, 0, 0), (0, 0, 0), 0.0, c)


def mat_diffuse_mirror(c, m):             # :156-158
    return material(c, c, (0, 0, 0), 0.0, m)


def sphere(center, radius, mat):          # :332-337  (radiusSquared = radius * radius in fp32)
    r = F32(radius)
    return np.concatenate([np.array(center, dtype=F32), [r], mat, [F32(r * r)]]).astype(F32)


def plane(center, normal, mat, tiled=True):   # :285-290 (isTiled is forced true by the reference)
    return np.concatenate([np.array(center, dtype=F32), np.array(normal, dtype=F32), mat, [1.0 if tiled else 0.0]]).astype(F32)


def light(pos, intensity):                # :251-254
    return np.array([*pos, intensity], dtype=F32)


@dataclass
class Scene:
    spheres: np.ndarray      # (ns, 18) float32
    planes: np.ndarray       # (np, 20) float32
    lights: np.ndarray       # (nl, 4) float32
    ambient: np.ndarray      # (3,) float32
    name: str = ""

    def __post_init__(self):
        self.spheres = np.ascontiguousarray(self.spheres, dtype=F32).reshape(-1, 18)
        self.planes = np.ascontiguousarray(self.planes, dtype=F32).reshape(-1, 20)
        self.lights = np.ascontiguousarray(self.lights, dtype=F32).reshape(-1, 4)
        self.ambient = np.ascontiguousarray(self.ambient, dtype=F32).reshape(3)


REF_AMBIENT = np.full(3, F32(43.0) / F32(255.0), dtype=F32)          # :469


def reference_plane():                    # :459-464
    return plane((0, -1.0, 0), (0, 1, 0), material((1, 1, 1), (0.5, 0.5, 0.5), (1, 1, 1), 0.5, (1, 1, 1)))


def default_scene() -> Scene:
    """The reference's hard-coded scene, RayTracer.cs:441-469."""
    spheres = np.stack([
        sphere((2.5, 0, 8), 1.0, mat_diffuse((1, 0, 0))),            # :442
        sphere((3, 0, 5), 1.0, mat_plastic((0, 1, 0))),              # :443
        sphere((-3, 1, 8), 1.0, mat_mirror((1, 1, 1))),              # :444
    ])
    lights = np.stack([light((-3, 1, -3), 1.0), light((33, 1, 10), 1.0)])   # :450-453
    return Scene(spheres, reference_plane()[None, :], lights, REF_AMBIENT.copy(), "default")


# --------------------------------------------------------------------------------------------------------
# Camera — RayTracer.cs:494-523 (basis, f64 trig cast to f32) and :892-896 (view-plane size)
# --------------------------------------------------------------------------------------------------------
NEAR_CLIP = F32(0.3)          # :481
FIELD_OF_VIEW = F32(60.0)     # :486
REFLECTION_RECURSION_LIMIT = 32   # :490


def camera_basis(yaw: float, pitch: float):
    yaw = float(F32(yaw)); pitch = float(F32(pitch))                # _yaw/_pitch are float fields :498-502
    fwd = np.array([F32(math.cos(pitch) * math.sin(yaw)), F32(-math.sin(pitch)), F32(math.cos(pitch) * math.cos(yaw))], dtype=F32)   # :511-513
    right = np.array([F32(math.cos(yaw)), F32(0.0), F32(-math.sin(yaw))], dtype=F32)                                              # :517-518
    l, r = right, fwd                                                 # Vector3.Cross(Right, Forward) :522-523
    up = np.array([F32(F32(l[1] * r[2]) - F32(l[2] * r[1])), F32(F32(l[2] * r[0]) - F32(l[0] * r[2])), F32(F32(l[0] * r[1]) - F32(l[1] * r[0]))], dtype=F32)
    return right, up, fwd


def view_params(width: int, height: int):
    deg2rad = F32(F32(math.pi) / F32(180.0))                         # MathHelper.DegreesToRadians (OpenTK)
    half = F32(F32(FIELD_OF_VIEW * F32(0.5)) * deg2rad)
    plane_h = F32(F32(NEAR_CLIP * F32(math.tan(float(half)))) * F32(2.0))      # :892
    aspect = F32(F32(width) / F32(height))                            # :893
    plane_w = F32(plane_h * aspect)                                   # :894
    return np.array([plane_w, plane_h, NEAR_CLIP], dtype=F32)         # :896


def make_camera(pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0, width=1280, height=720) -> np.ndarray:
    """15 floats: pos, right, up, forward, view_params — the rt_camera struct."""
    right, up, fwd = camera_basis(yaw, pitch)
    return np.concatenate([np.array(pos, dtype=F32), right, up, fwd, view_params(width, height)]).astype(F32)


# --------------------------------------------------------------------------------------------------------
# PCG32 (XSH-RR 64/32, O'Neill), stream 1 — the generator SURVEY §8d names for the synthetic scenes
# --------------------------------------------------------------------------------------------------------
class PCG32:
    MULT = 6364136223846793005
    MASK = (1 << 64) - 1

    def __init__(self, seed: int, stream: int = 1):
        self.inc = ((stream << 1) | 1) & self.MASK
        self.state = 0
        self.next_u32()
        self.state = (self.state + seed) & self.MASK
        self.next_u32()

    def next_u32(self) -> int:
        old = self.state
        self.state = (old * self.MULT + self.inc) & self.MASK
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def u(self) -> np.float32:
        """uniform [0,1) with 24 bits, exactly representable in fp32"""
        return F32(self.next_u32() >> 8) * F32(5.9604644775390625e-08)


def random_spheres_scene(n: int, seed: int, x_half: float, z_lo: float, z_hi: float, name: str) -> Scene:
    """SURVEY §8d configs 3/4.  Per sphere, draws in this order (all arithmetic fp32):
         x = -x_half + 2*x_half*u;  z = z_lo + (z_hi-z_lo)*u;  r = 0.15 + 0.45*u;  y = -1 + r + 2*u*u;
         k = u (material class);  cr,cg,cb = u,u,u;  s = u (specularity choice)
       classes: k<0.6 Diffuse(rgb) | k<0.8 Plastic(rgb, n in {1,8,32}) | k<0.9 Metal(rgb, n in {4,16}) | Mirror(.9,.9,.9)
    """
    rng = PCG32(seed, 1)
    out = np.zeros((n, 18), dtype=F32)
    xh, zl, zh = F32(x_half), F32(z_lo), F32(z_hi)
    for i in range(n):
        x = F32(-xh + F32(F32(2.0) * xh) * rng.u())
        z = F32(zl + F32(zh - zl) * rng.u())
        r = F32(F32(0.15) + F32(0.45) * rng.u())
        uy = rng.u()
        y = F32(F32(F32(-1.0) + r) + F32(F32(2.0) * F32(uy * uy)))
        k = rng.u()
        rgb = (rng.u(), rng.u(), rng.u())
        s = rng.u()
        if k < F32(0.6):
            m = mat_diffuse(rgb)
        elif k < F32(0.8):
            m = mat_plastic(rgb, (1.0, 8.0, 32.0)[min(int(s * F32(3.0)), 2)])
        elif k < F32(0.9):
            m = mat_metal(rgb, (4.0, 16.0)[min(int(s * F32(2.0)), 1)])
        else:
            m = mat_mirror((0.9, 0.9, 0.9))
        out[i] = sphere((x, y, z), r, m)
    return out, name


def config3_scene() -> Scene:
    """1,024 random spheres + reference ground plane + 4 point lights (BASELINE.json configs[2])."""
    sph, name = random_spheres_scene(1024, 1024, 24.0, 4.0, 52.0, "config3_1024")
    lights = np.stack([light((-20, 12, 10), 1.0), light((20, 12, 10), 1.0), light((-20, 12, 40), 1.0), light((20, 12, 40), 1.0)])
    return Scene(sph, reference_plane()[None, :], lights, REF_AMBIENT.copy(), name)


def config4_scene(n: int = 100_000) -> Scene:
    """100k random spheres (BASELINE.json configs[3]); x in [-150,150], z in [4,304]."""
    sph, name = random_spheres_scene(n, 100000, 150.0, 4.0, 304.0, f"config4_{n}")
    lights = np.stack([light((-20, 12, 10), 1.0), light((20, 12, 10), 1.0), light((-20, 12, 40), 1.0), light((20, 12, 40), 1.0)])
    return Scene(sph, reference_plane()[None, :], lights, REF_AMBIENT.copy(), name)


def small_random_scene(n: int, seed: int) -> Scene:
    """Small parity-test scenes: n spheres of every material class incl. DiffuseMirror, 3 lights, 2 planes."""
    sph, _ = random_spheres_scene(n, seed, 6.0, 3.0, 14.0, f"small_{n}_{seed}")
    rng = PCG32(seed ^ 0x5EED, 1)
    if n > 0:
        sph[0] = sphere((sph[0][0], sph[0][1], sph[0][2]), sph[0][3], mat_diffuse_mirror((rng.u(), rng.u(), rng.u()), (0.5, 0.6, 0.7)))
    planes = np.stack([
        reference_plane(),
        plane((0, 0, 18.0), (0.1, 0.05, -1.0), material((0.3, 0.5, 0.9), (0.2, 0.2, 0.2), (0.4, 0.4, 0.4), 8.0, (0, 0, 0))),
    ])
    lights = np.stack([light((-3, 4, -2), 1.0), light((6, 5, 2), 0.7), light((0, 8, 10), 0.5)])
    return Scene(sph, planes, lights, REF_AMBIENT.copy(), f"small_{n}_{seed}")


SCALED_CAMERA = dict(pos=(0.0, 3.0, -6.0), yaw=0.0, pitch=0.25)      # configs 3/4 camera


def degenerate_scene() -> Scene:
    """Parity edge cases the reference handles "by accident of IEEE": zero / negative radiusSquared, a light at the origin
    (shadow direction = zero vector, RayTracer.cs:574), a light inside a sphere, a plane with a zero normal (0/0 = NaN: never
    hit), a non-unit plane normal (used as given, :743), negative and > 1 colours, tiny and large specular exponents, a sphere
    far away (fp32 noise regime), overlapping spheres < 0.01 apart (order-dependent secondary fold, :804)."""
    s = [
        sphere((0.0, 0.0, 6.0), 1.0, mat_mirror((1.2, 0.9, 0.8))),
        sphere((0.004, 0.0, 6.003), 1.0, mat_plastic((0.2, 0.7, 1.5), 64.0)),          # overlaps the mirror within 0.01
        sphere((-2.5, 0.2, 5.0), 0.8, mat_metal((0.9, -0.3, 0.4), 0.001)),             # negative colour, tiny exponent
        sphere((2.5, 0.0, 5.0), 0.0, mat_diffuse((1, 1, 1))),                           # radius 0
        sphere((1.5, 1.5, 4.0), 0.5, mat_diffuse_mirror((0.3, 0.3, 0.3), (0.5, 0.5, 0.5))),
        sphere((400.0, 30.0, 900.0), 0.3, mat_diffuse((1, 1, 0))),                      # far: discriminant noise > r^2
        sphere((-1.0, 3.0, 7.0), 0.7, mat_plastic((0.5, 0.5, 0.5), 3.5)),               # general Math.Pow exponent
    ]
    s = np.stack(s)
    s[3, 17] = -0.25                                                                    # negative radiusSquared
    planes = np.stack([
        reference_plane(),
        plane((0, 0, 0), (0, 0, 0), mat_diffuse((1, 1, 1))),                            # zero normal
        plane((0, 0, 30.0), (0.0, 0.5, -3.0), material((0.4, 0.4, 0.9), (0.3, 0.3, 0.3), (0.2, 0.2, 0.2), 2.0, (0.3, 0.3, 0.3))),
    ])
    lights = np.stack([light((0, 0, 0), 1.0), light((0.0, 0.1, 6.0), 2.0), light((-4, 5, 0), 1.0), light((0, 8, 10), 0.5)])
    return Scene(s, planes, lights, REF_AMBIENT.copy(), "degenerate")
