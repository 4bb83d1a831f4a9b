"""Second, independent restatement of Raytracer/RayTracer.cs in vectorised numpy fp32 — a cross-check of the C++ oracle.

Written directly from the C# source (not from oracle/rt_oracle.cpp) and in a different style (whole-frame array operations,
no recursion): `render_cap0` evaluates the frame for ReflectionRecursionLimit = 0, where the recursion is exactly one level deep —
a primary hit is fully shaded (shadow rays, Phong, checkerboard, ambient) and its mirror term is the TERMINAL colour of the
secondary hit (bounce 1 > 0: plane -> white :734, sphere -> black :843, nothing / too-close hit -> black). That exercises
every arithmetic expression of the path: primary-ray generation (:963-971), IntersectsSphere (:613-642), IntersectPlane
(:590-604), both folds (:975-993, :792-825), IntersectShadowLight (:573-582), ShapePhongShading (:665-695), the checkerboard
(:756-771), attenuation (:754, :866) and ShiftColor (:1046-1052).  numpy float32 element-wise ops are single IEEE operations
(no FMA contraction), `np.float64` is used exactly where the C# promotes to double.
TEST INFRASTRUCTURE ONLY."""
import numpy as np

F = np.float32


def _dot(a, b):                      # OpenTK Vector3.Dot: (x*x') + (y*y') + (z*z')
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _normalize(v):                   # OpenTK Vector3.Normalize: scale = 1f / Length; v * scale
    length = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    with np.errstate(divide="ignore", invalid="ignore"):
        s = F(1.0) / length
        return [v[0] * s, v[1] * s, v[2] * s]


def _cs_max0(x):                     # Math.Max(0, x) / Math.Max(x, 0): NaN-propagating
    return np.where(x <= 0, F(0.0), x).astype(F)


def _intersect_sphere(o, d, c, r2, eps):
    """IntersectsSphere :613-642 literally (both roots, Min/Max), vectorised. Returns (collision, distance)."""
    oc = [o[0] - c[0], o[1] - c[1], o[2] - c[2]]
    a = _dot(d, d)
    b = F(2) * _dot(oc, d)
    cc = _dot(oc, oc) - r2
    disc = b * b - F(4) * a * cc
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(np.where(disc >= 0, disc, F(0)).astype(np.float64)).astype(F)       # (float)Math.Sqrt(d)
        a2 = F(2) * a
        dist2 = (-b + sq) / a2
        dist1 = (-b - sq) / a2
        d1e, d2e = dist1 - F(eps), dist2 - F(eps)
        mx = lambda x: np.where(np.isnan(x), x, np.maximum(x, F(0)))                       # Math.Max(x, 0)
        mn = lambda x, y: np.where(np.isnan(x) | np.isnan(y), F(np.nan), np.minimum(x, y)) # Math.Min
        distance = mn(mx(dist1), mx(dist2))
        distance_eps = mn(mx(d1e), mx(d2e))
    hit = (disc >= 0) & (distance_eps > 0)
    return hit, np.where(hit, distance, F(0)).astype(F)


def _intersect_plane(o, d, pc, pn):
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (-o[0] * pn[0] - o[1] * pn[1] - o[2] * pn[2] + _dot(pc, pn)) / _dot(d, pn)      # :591-596
    hit = t > 0
    return hit, np.where(hit, t, F(0)).astype(F)


def _shift_color(c):
    ch = []
    for k in range(3):
        v = np.where(c[k] < 0, F(0), np.where(c[k] > 1, F(1), c[k])).astype(F) * F(255)    # Math.Clamp * 255f
        vi = np.where(np.isnan(v), 0, np.floor(v.astype(np.float64))).astype(np.int64) & 255
        ch.append(vi)
    return ((ch[0] << 16) | (ch[1] << 8) | ch[2]).astype(np.int32)


def render_cap0(scene, cam15, w, h):
    """Frame for ReflectionRecursionLimit = 0. Returns (pixels int32[h,w], primary hit code int32[h,w], primary t f32[h,w])."""
    sph, pls, lts, amb = scene.spheres, scene.planes, scene.lights, scene.ambient
    ns, npl = len(sph), len(pls)
    cam15 = np.asarray(cam15, F)
    pos, right, up, fwd, view = cam15[0:3], cam15[3:6], cam15[6:9], cam15[9:12], cam15[12:15]
    ys, xs = np.meshgrid(np.arange(h, dtype=F), np.arange(w, dtype=F), indexing="ij")
    u = xs / F(w) - F(0.5)                                                                # :964
    v = ys / F(h) - F(0.5)
    lx, ly, lz = u * view[0], v * view[1], np.full_like(u, F(1.0) * view[2])               # :965
    vp = [((pos[k] + right[k] * lx) + up[k] * ly) + fwd[k] * lz for k in range(3)]          # :967-969
    d0 = _normalize([vp[k] - pos[k] for k in range(3)])                                    # :971
    o0 = [np.full_like(u, pos[k]) for k in range(3)]

    def mat(rec13):
        return dict(kd=rec13[0:3], ka=rec13[3:6], ks=rec13[6:9], n=rec13[9], km=rec13[10:13])

    def folds(o, d, secondary):
        """Returns per-pixel (is_sphere, index, distance, any_hit) for the primary (:973-993) or secondary (:789-825) fold."""
        best_s = np.full(u.shape, F(np.inf)); idx_s = np.full(u.shape, -1)
        for i in range(ns):
            hit, dist = _intersect_sphere(o, d, sph[i, 0:3], sph[i, 17], 0.0)
            if secondary:
                ok = (dist - F(0.01) > 0) & (dist - F(0.01) < best_s)                      # :804
            else:
                ok = (dist > 0) & (best_s > dist)                                         # :977
            best_s = np.where(ok, dist, best_s).astype(F); idx_s = np.where(ok, i, idx_s)
        best_p = np.full(u.shape, F(np.inf)); idx_p = np.full(u.shape, -1)
        for i in range(npl):
            hit, dist = _intersect_plane(o, d, pls[i, 0:3], pls[i, 3:6])
            ok = (dist > 0) & (best_p > dist)                                             # :987 / :819
            best_p = np.where(ok, dist, best_p).astype(F); idx_p = np.where(ok, i, idx_p)
        is_s = best_s < best_p                                                            # :993 / :825
        return is_s, np.where(is_s, idx_s, idx_p), np.where(is_s, best_s, best_p).astype(F), is_s | (idx_p >= 0)

    is_s, idx, dist, anyhit = folds(o0, d0, False)
    code = np.where(~anyhit, -1, np.where(is_s, idx, ns + idx)).astype(np.int32)
    tsel = np.where(anyhit, dist, F(0)).astype(F)

    col = [np.zeros(u.shape, F) for _ in range(3)]
    shaded = anyhit & ~(dist - F(0.01) <= 0)                                              # :839 / :731 (bounce 0 <= cap)
    hitp = [o0[k] + d0[k] * dist for k in range(3)]                                        # :846 / :736

    prims = [("s", i) for i in range(ns)] + [("p", i) for i in range(npl)]
    for kind, i in prims:
        sel = shaded & (is_s if kind == "s" else ~is_s) & (idx == i)
        if not sel.any():
            continue
        m = mat(sph[i, 4:17] if kind == "s" else pls[i, 6:19])
        if kind == "s":
            N = _normalize([hitp[k] - sph[i, k] for k in range(3)])                        # :706 / :854
        else:
            N = [np.full(u.shape, pls[i, 3 + k], F) for k in range(3)]
        c = [np.zeros(u.shape, F) for _ in range(3)]
        if np.any(m["km"] != 0):                                                          # IsMirror :85
            s2 = F(2) * _dot(d0, N)
            rd = [d0[k] - s2 * N[k] for k in range(3)]                                     # :719
            is_s2, idx2, dist2, any2 = folds(hitp, rd, True)
            # bounce = 1 > cap = 0: the secondary hit returns its terminal colour — unless it is too close (black)
            term = np.where(any2 & ~(dist2 - F(0.01) <= 0) & ~is_s2, F(1.0), F(0.0)).astype(F)   # plane white :734, sphere black :843
            c = [c[k] + term * m["km"][k] for k in range(3)]                               # :857 / :746
        if np.any(m["kd"] != 0):                                                          # IsDiffuse :89
            for li in range(len(lts)):
                lp, inten = lts[li, 0:3], lts[li, 3]
                occluded = np.zeros(u.shape, bool)
                ldir = [np.full(u.shape, lp[k], F) for k in range(3)]                      # direction = light POSITION :574
                for j in range(ns):
                    hj, _ = _intersect_sphere(hitp, ldir, sph[j, 0:3], sph[j, 17], 0.001)
                    occluded |= hj
                I = np.where(occluded, F(0), inten).astype(F)                             # :581
                L = _normalize([lp[k] - hitp[k] for k in range(3)])                        # :667
                V = _normalize(d0)                                                        # :668
                ph = [m["kd"][k] * _cs_max0(_dot(N, L)) for k in range(3)]                 # :672-678
                if np.any(m["ks"] != 0) and m["n"] > 0:                                   # HasSpecularity :93
                    s2 = F(2) * _dot(L, N)
                    rv = [L[k] - s2 * N[k] for k in range(3)]                              # :683-684
                    sp = _dot(V, _normalize(rv))
                    with np.errstate(invalid="ignore"):
                        pw = np.power(_cs_max0(sp).astype(np.float64), np.float64(m["n"])).astype(F)   # :691
                    ph = [ph[k] + m["ks"][k] * pw for k in range(3)]
                if kind == "s":
                    with np.errstate(divide="ignore", invalid="ignore"):
                        att = F(1) / dist * dist                                          # :866
                    term3 = [(I * att) * ph[k] for k in range(3)]                          # :868-869
                else:
                    with np.errstate(divide="ignore"):
                        att = (1.0 / np.power(dist.astype(np.float64), 2.0)).astype(F)     # :754
                    pn = [F(pls[i, 3]), F(pls[i, 4]), F(pls[i, 5])]
                    cr = [pn[1] * F(0) - pn[2] * F(0), pn[2] * F(1) - pn[0] * F(0), pn[0] * F(0) - pn[1] * F(1)]   # Cross(n,(1,0,0)) :760
                    e1 = _normalize([np.float32(x) for x in cr])
                    e2c = [pn[1] * e1[2] - pn[2] * e1[1], pn[2] * e1[0] - pn[0] * e1[2], pn[0] * e1[1] - pn[1] * e1[0]]
                    e2 = _normalize([np.float32(x) for x in e2c])                          # :765
                    uu = _dot(e1, hitp); vv = _dot(e2, hitp)                               # :766-767
                    with np.errstate(invalid="ignore"):
                        cb = (np.trunc(uu).astype(np.int64) + np.trunc(vv).astype(np.int64)) & 1     # :769
                    tile = cb.astype(F)
                    term3 = [np.where(np.isnan(t), t, np.maximum(t, F(0))).astype(F)
                             for t in [((I * att) * ph[k]) * tile for k in range(3)]]      # :774-775
                c = [c[k] + term3[k] for k in range(3)]
        c = [c[k] + amb[k] * m["ka"][k] for k in range(3)]                                 # :873 / :778
        for k in range(3):
            col[k] = np.where(sel, c[k], col[k]).astype(F)
    return _shift_color(col), code, tsel


# ---------------------------------------------------------------------------------------------------------------------------
# Any recursion limit: the same restatement with the recursion of TraceSecondaryRay (:789-826) followed on index subsets.
# Written after render_cap0 (which stays as it is: whole-frame masks, one level) — here every ray set is a flat array and a
# mirror hit recurses on the rays that hit that primitive.
# ---------------------------------------------------------------------------------------------------------------------------
def _folds_flat(scene, o, d, secondary):
    sph, pls = scene.spheres, scene.planes
    n = len(o[0])
    best_s = np.full(n, F(np.inf)); idx_s = np.full(n, -1)
    for i in range(len(sph)):
        _, dist = _intersect_sphere(o, d, sph[i, 0:3], sph[i, 17], 0.0)
        if secondary:
            ok = (dist - F(0.01) > 0) & (dist - F(0.01) < best_s)                          # :804
        else:
            ok = (dist > 0) & (best_s > dist)                                             # :977
        best_s = np.where(ok, dist, best_s).astype(F); idx_s = np.where(ok, i, idx_s)
    best_p = np.full(n, F(np.inf)); idx_p = np.full(n, -1)
    for i in range(len(pls)):
        _, dist = _intersect_plane(o, d, pls[i, 0:3], pls[i, 3:6])
        ok = (dist > 0) & (best_p > dist)                                                 # :987 / :819
        best_p = np.where(ok, dist, best_p).astype(F); idx_p = np.where(ok, i, idx_p)
    is_s = best_s < best_p                                                                # :993 / :825
    return is_s, np.where(is_s, idx_s, idx_p), np.where(is_s, best_s, best_p).astype(F), is_s | (idx_p >= 0)


def _trace_flat(scene, o, d, bounce, cap, secondary):
    """Colour (3 float32 arrays) of the rays (o, d); also (hit code, selected distance) for the caller's bookkeeping."""
    sph, pls, lts, amb = scene.spheres, scene.planes, scene.lights, scene.ambient
    ns = len(sph)
    n = len(o[0])
    is_s, idx, dist, anyhit = _folds_flat(scene, o, d, secondary)
    code = np.where(~anyhit, -1, np.where(is_s, idx, ns + idx)).astype(np.int32)
    tsel = np.where(anyhit, dist, F(0)).astype(F)
    col = [np.zeros(n, F) for _ in range(3)]
    live = anyhit & ~(dist - F(0.01) <= 0)                                                # :839 / :731
    if bounce > cap:                                                                      # :843 sphere black / :734 plane white
        white = live & ~is_s
        return [np.where(white, F(1.0), F(0.0)).astype(F) for _ in range(3)], code, tsel
    prims = [("s", i) for i in range(ns)] + [("p", i) for i in range(len(pls))]
    for kind, i in prims:
        sub = np.flatnonzero(live & (is_s if kind == "s" else ~is_s) & (idx == i))
        if not len(sub):
            continue
        os_ = [o[k][sub] for k in range(3)]; ds = [d[k][sub] for k in range(3)]; dd = dist[sub]
        hitp = [os_[k] + ds[k] * dd for k in range(3)]                                     # :846 / :736
        rec = sph[i, 4:17] if kind == "s" else pls[i, 6:19]
        kd, ka, ks, nspec, km = rec[0:3], rec[3:6], rec[6:9], rec[9], rec[10:13]
        if kind == "s":
            N = _normalize([hitp[k] - sph[i, k] for k in range(3)])                        # :706 / :854
        else:
            N = [np.full(len(sub), pls[i, 3 + k], F) for k in range(3)]
        c = [np.zeros(len(sub), F) for _ in range(3)]
        if np.any(km != 0):                                                               # IsMirror :85
            s2 = F(2) * _dot(ds, N)
            rd = [ds[k] - s2 * N[k] for k in range(3)]                                     # :719
            cin, _, _ = _trace_flat(scene, hitp, rd, bounce + 1, cap, True)               # :857 / :746 (bounce incremented first)
            c = [c[k] + cin[k] * km[k] for k in range(3)]
        if np.any(kd != 0):                                                               # IsDiffuse :89
            for li in range(len(lts)):
                lp, inten = lts[li, 0:3], lts[li, 3]
                occluded = np.zeros(len(sub), bool)
                ldir = [np.full(len(sub), lp[k], F) for k in range(3)]                     # direction = light POSITION :574
                for j in range(ns):
                    hj, _ = _intersect_sphere(hitp, ldir, sph[j, 0:3], sph[j, 17], 0.001)
                    occluded |= hj
                I = np.where(occluded, F(0), inten).astype(F)                             # :581
                L = _normalize([lp[k] - hitp[k] for k in range(3)])                        # :667
                V = _normalize(ds)                                                        # :668
                ph = [kd[k] * _cs_max0(_dot(N, L)) for k in range(3)]                      # :672-678
                if np.any(ks != 0) and nspec > 0:                                         # HasSpecularity :93
                    s2 = F(2) * _dot(L, N)
                    rv = [L[k] - s2 * N[k] for k in range(3)]                              # :683-684
                    sp = _dot(V, _normalize(rv))
                    with np.errstate(invalid="ignore"):
                        pw = np.power(_cs_max0(sp).astype(np.float64), np.float64(nspec)).astype(F)   # :691
                    ph = [ph[k] + ks[k] * pw for k in range(3)]
                if kind == "s":
                    with np.errstate(divide="ignore", invalid="ignore"):
                        att = F(1) / dd * dd                                              # :866
                    term3 = [(I * att) * ph[k] for k in range(3)]                          # :868-869
                else:
                    with np.errstate(divide="ignore"):
                        att = (1.0 / np.power(dd.astype(np.float64), 2.0)).astype(F)       # :754
                    pn = [F(pls[i, 3]), F(pls[i, 4]), F(pls[i, 5])]
                    cr = [pn[1] * F(0) - pn[2] * F(0), pn[2] * F(1) - pn[0] * F(0), pn[0] * F(0) - pn[1] * F(1)]   # :760
                    e1 = _normalize([np.float32(x) for x in cr])
                    if e1[0] == 0 and e1[1] == 0 and e1[2] == 0:                           # :761-762
                        cr = [pn[1] * F(1) - pn[2] * F(0), pn[2] * F(0) - pn[0] * F(1), pn[0] * F(0) - pn[1] * F(0)]
                        e1 = _normalize([np.float32(x) for x in cr])
                    e2c = [pn[1] * e1[2] - pn[2] * e1[1], pn[2] * e1[0] - pn[0] * e1[2], pn[0] * e1[1] - pn[1] * e1[0]]
                    e2 = _normalize([np.float32(x) for x in e2c])                          # :765
                    uu = _dot(e1, hitp); vv = _dot(e2, hitp)                               # :766-767
                    with np.errstate(invalid="ignore"):
                        cb = (np.trunc(uu).astype(np.int64) + np.trunc(vv).astype(np.int64)) & 1     # :769
                    tile = cb.astype(F)
                    term3 = [np.where(np.isnan(t), t, np.maximum(t, F(0))).astype(F)
                             for t in [((I * att) * ph[k]) * tile for k in range(3)]]      # :774-775
                c = [c[k] + term3[k] for k in range(3)]
        c = [c[k] + amb[k] * ka[k] for k in range(3)]                                      # :873 / :778
        for k in range(3):
            col[k][sub] = c[k]
    return col, code, tsel


def render(scene, cam15, w, h, cap):
    """Frame for any ReflectionRecursionLimit. Returns (pixels int32[h,w], primary hit code int32[h,w], primary t f32[h,w])."""
    cam15 = np.asarray(cam15, F)
    pos, right, up, fwd, view = cam15[0:3], cam15[3:6], cam15[6:9], cam15[9:12], cam15[12:15]
    ys, xs = np.meshgrid(np.arange(h, dtype=F), np.arange(w, dtype=F), indexing="ij")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    u = xs / F(w) - F(0.5)                                                                # :964
    v = ys / F(h) - F(0.5)
    lx, ly, lz = u * view[0], v * view[1], np.full_like(u, F(1.0) * view[2])               # :965
    vp = [((pos[k] + right[k] * lx) + up[k] * ly) + fwd[k] * lz for k in range(3)]          # :967-969
    d0 = _normalize([vp[k] - pos[k] for k in range(3)])                                    # :971
    o0 = [np.full_like(u, pos[k]) for k in range(3)]
    col, code, tsel = _trace_flat(scene, o0, d0, 0, cap, False)
    return _shift_color(col).reshape(h, w), code.reshape(h, w), tsel.reshape(h, w)
