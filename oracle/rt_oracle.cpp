// rt_oracle.cpp — CPU ORACLE for the per-pixel ray-cast path of Raytracer/RayTracer.cs.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.  The product (librtb200.so) never links or calls it.
//
// PARITY UNPINNED: the reference (C#/.NET 6) cannot be executed in this image (no dotnet/mono/csc) and holds
// no tests, golden vectors or fixtures for this path.  Vector arithmetic lives in NuGet OpenTK 4.7.1
// (Raytracer/InfogrRaytracer.csproj:12), not vendored; its published semantics are restated below
// (Dot = (x*x')+(y*y')+(z*z'); Normalize = v * (1f/sqrt(x*x+y*y+z*z)); Cross standard; V3*V3 component-wise).
// Build with -DORC_NORMALIZE_TRUE_DIV to switch Normalize to true division for a sensitivity diff.
//
// Evaluation model restated: .NET 6 x64 RyuJIT -> scalar SSE fp32, round-to-nearest-even, no FMA contraction,
// denormals on, left-to-right association as written; Math.X(double) sites promote to f64 and cast back.
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp (see oracle/Makefile). Never -march=native/-ffast-math.
//
// Every function cites the reference file:line it follows (all in Raytracer/RayTracer.cs unless noted).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <climits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------------------
// BCL restatements (System.Math, .NET 6)
// ---------------------------------------------------------------------------------------------------------
static inline bool is_neg(float v) { uint32_t b; std::memcpy(&b, &v, 4); return (b >> 31) != 0; }

// Math.Max(float,float): IEEE 754-2019 maximum, NaN-propagating, +0 > -0.
static inline float cs_max(float a, float b) {
    if (a != b) {
        if (!std::isnan(a)) return b < a ? a : b;
        return a;
    }
    return is_neg(b) ? a : b;
}
// Math.Min(float,float): IEEE 754-2019 minimum.
static inline float cs_min(float a, float b) {
    if (a != b) {
        if (!std::isnan(a)) return a < b ? a : b;
        return a;
    }
    return is_neg(a) ? a : b;
}
// Math.Clamp(float,float,float)
static inline float cs_clamp(float v, float lo, float hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
// (int)float on x64 RyuJIT (.NET 6): cvttss2si -> 0x80000000 for NaN / out of range.
static inline int32_t cs_f2i(float f) {
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT_MIN;
    return (int32_t)f;
}
// (int)double on x64 RyuJIT (.NET 6): cvttsd2si.
static inline int32_t cs_d2i(double d) {
    if (!(d > -2147483649.0 && d < 2147483648.0)) return INT_MIN;
    return (int32_t)d;
}

// ---------------------------------------------------------------------------------------------------------
// OpenTK.Mathematics.Vector3 restatement (OpenTK 4.7.1, not in tree — see header)
// ---------------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static inline V3 splat(float f) { return v3(f, f, f); }                                   // VecUtil.FromFloat3 :20
static inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 mulf(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 mulv(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
static inline V3 cross(V3 l, V3 r) {
    return v3((l.y * r.z) - (l.z * r.y), (l.z * r.x) - (l.x * r.z), (l.x * r.y) - (l.y * r.x));
}
static inline V3 normalize(V3 v) {
    float len = std::sqrt((v.x * v.x) + (v.y * v.y) + (v.z * v.z));
#ifdef ORC_NORMALIZE_TRUE_DIV
    return v3(v.x / len, v.y / len, v.z / len);
#else
    float scale = 1.0f / len;
    return v3(v.x * scale, v.y * scale, v.z * scale);
#endif
}
static inline bool is_zero(V3 v) { return v.x == 0 && v.y == 0 && v.z == 0; }              // VecUtil.IsZero :52
static inline V3 vmaxf(V3 a, float f) { return v3(cs_max(a.x, f), cs_max(a.y, f), cs_max(a.z, f)); } // VecUtil.Max :39

// ---------------------------------------------------------------------------------------------------------
// Scene value types — flat float records in the C# field order (the C-ABI layout, include/rtb200.h)
// ---------------------------------------------------------------------------------------------------------
struct Material {                 // :60-80
    V3 kd, ka, ks; float n; V3 km;
    bool is_mirror() const { return !is_zero(km); }                                       // :85
    bool is_diffuse() const { return !is_zero(kd); }                                      // :89
    bool has_spec() const { return !is_zero(ks) && n > 0.0f; }                            // :93
};
struct Sphere { V3 c; float r; Material m; float r2; };                                   // :308-338 (18 floats)
struct Plane { V3 c; V3 n; Material m; bool tiled; };                                     // :260-303 (20 floats)
struct Light { V3 p; float i; };                                                          // :236-255 (4 floats)
struct Ray { V3 o, d; };                                                                  // :366-421 (rayKind is debug-only)

static Material load_mat(const float* f) {
    Material m; m.kd = v3(f[0], f[1], f[2]); m.ka = v3(f[3], f[4], f[5]); m.ks = v3(f[6], f[7], f[8]);
    m.n = f[9]; m.km = v3(f[10], f[11], f[12]); return m;
}

struct Counters {
    // nearest-first deterministic accounting: rays on the SELECTED hit chain only
    uint64_t primary = 0, shadow = 0, secondary = 0;
    // flop-model inputs: sphere tests, sphere tests with discriminant >= 0, plane tests, shaded hits
    uint64_t sphere_tests = 0, sphere_disc_pos = 0, plane_tests = 0;
    uint64_t shade_diffuse = 0, shade_specular = 0, shade_mirror = 0, shaded_hits = 0;
    // faithful mode only: every ray the reference actually builds (it shades non-nearest hits too)
    uint64_t faithful_secondary = 0, faithful_shadow = 0;
    void add(const Counters& o) {
        primary += o.primary; shadow += o.shadow; secondary += o.secondary;
        sphere_tests += o.sphere_tests; sphere_disc_pos += o.sphere_disc_pos; plane_tests += o.plane_tests;
        shade_diffuse += o.shade_diffuse; shade_specular += o.shade_specular; shade_mirror += o.shade_mirror;
        shaded_hits += o.shaded_hits; faithful_secondary += o.faithful_secondary; faithful_shadow += o.faithful_shadow;
    }
};
static const int N_COUNTERS = 12;

struct Scene {
    std::vector<Sphere> spheres; std::vector<Plane> planes; std::vector<Light> lights;
    V3 ambient; int cap;
};

// Order-independent chain hash (debug AOV shared with the CUDA path): wrapping sum of per-event hashes.
static inline uint32_t mix32(uint32_t h, uint32_t v) { h ^= v; h *= 16777619u; h ^= h >> 15; return h; }
static inline uint32_t event_hash(uint32_t level, uint32_t kind, uint32_t a, uint32_t b) {
    uint32_t h = 0x811C9DC5u;
    h = mix32(h, level); h = mix32(h, kind); h = mix32(h, a); h = mix32(h, b);
    return h;
}
static inline uint32_t fbits(float f) { uint32_t b; std::memcpy(&b, &f, 4); return b; }

// One record per ray of the nearest-first chain: the per-ray digest of the reference's DEBUG_ENABLE TracedRay log (:424-435,
// filled per intersection test at :601, :639, :801). Same layout as rt_ray_record (include/rtb200.h).
struct RayRecord {
    float origin[3], direction[3], hit_point[3], distance;
    int32_t hit; uint32_t kind, pixel, level, light, reserved;
};
static_assert(sizeof(RayRecord) == 64, "RayRecord layout");

struct Ctx {            // per-thread trace context
    std::vector<RayRecord>* log = nullptr; uint32_t log_pixel = 0;    // orc_ray_log only
    const Scene* sc; Counters cnt; uint32_t hash;
    bool faithful;      // true: shade-all-then-select exactly as written; false: select-then-shade (A.11)
    bool count_chain;   // == !faithful: chain counters / hash are only meaningful in nearest-first mode
};

struct Isect { bool hit; float d; };                                                      // IntersectResult :164-201

// IntersectsSphere :613-642
static Isect intersects_sphere(Ctx& cx, const Ray& ray, const Sphere& s, float eps) {
    V3 oc = sub(ray.o, s.c);                                                              // :614
    Isect res = {false, 0.0f};                                                            // :615 Miss()
    float a = dot(ray.d, ray.d);                                                          // :617
    float b = 2 * dot(oc, ray.d);                                                         // :618
    float c = dot(oc, oc) - s.r2;                                                         // :619
    float d = b * b - 4 * a * c;                                                          // :621  (4*a)*c
    if (cx.count_chain) cx.cnt.sphere_tests++;
    if (d >= 0) {
        if (cx.count_chain) cx.cnt.sphere_disc_pos++;
        float dsqrt = (float)std::sqrt((double)d);                                        // :623
        float a2 = 2 * a;                                                                 // :624
        float dist2 = (-b + dsqrt) / a2;                                                  // :626
        float dist1 = (-b - dsqrt) / a2;                                                  // :627
        float d1e = dist1 - eps;                                                          // :629
        float d2e = dist2 - eps;                                                          // :630
        float dist = cs_min(cs_max(dist1, 0), cs_max(dist2, 0));                          // :632
        float diste = cs_min(cs_max(d1e, 0), cs_max(d2e, 0));                             // :633
        if (diste > 0) { res.hit = true; res.d = dist; }                                  // :635
    }
    return res;
}

// IntersectPlane :590-604
static Isect intersect_plane(Ctx& cx, const Ray& ray, const Plane& p) {
    float t = (-ray.o.x * p.n.x - ray.o.y * p.n.y - ray.o.z * p.n.z + dot(p.c, p.n))      // :591-594
              / dot(ray.d, p.n);                                                          // :596
    if (cx.count_chain) cx.cnt.plane_tests++;
    Isect r; r.hit = t > 0; r.d = r.hit ? t : 0.0f;                                       // :598
    return r;
}

static void log_ray(Ctx& cx, uint32_t kind, uint32_t level, uint32_t light, int32_t hit, V3 o, V3 d, float dist) {
    if (!cx.log) return;
    V3 hp = add(o, mulf(d, dist));                                                        // TracedRay.hitPoint :601 / :639
    RayRecord r = {{o.x, o.y, o.z}, {d.x, d.y, d.z}, {hp.x, hp.y, hp.z}, dist, hit, kind, cx.log_pixel, level, light, 0u};
    cx.log->push_back(r);
}
// Shadow-ray record: the nearest colliding sphere of IntersectShadowLight's loop (:577-579), lowest index on ties.
static void log_shadow(Ctx& cx, uint32_t level, uint32_t li, V3 hit, const Light& l) {
    if (!cx.log) return;
    const bool cc = cx.count_chain; cx.count_chain = false;                               // do not disturb the counters
    Ray ray = {hit, l.p};
    int sel = -1; float best = INFINITY;
    for (int i = 0; i < (int)cx.sc->spheres.size(); i++) {
        Isect is = intersects_sphere(cx, ray, cx.sc->spheres[i], 0.001f);
        if (is.hit && is.d < best) { best = is.d; sel = i; }
    }
    cx.count_chain = cc;
    log_ray(cx, 2, level, li, sel, hit, l.p, sel >= 0 ? best : 0.0f);
}

// IntersectShadowLight :573-582 — direction is the light POSITION, spheres only, no early-out, unbounded t.
static float intersect_shadow_light(Ctx& cx, V3 hit, const Light& l) {
    Ray ray = {hit, l.p};                                                                 // :574
    bool obstructed = false;
    for (const Sphere& s : cx.sc->spheres)                                                // :577
        if (intersects_sphere(cx, ray, s, 0.001f).hit) obstructed = true;                 // :578
    return obstructed ? 0.0f : l.i;                                                       // :581
}

// ShapePhongShading :665-695
static V3 shape_phong(Ctx& cx, V3 hit, const Ray& ray, V3 n, const Material& m, const Light& l) {
    V3 L = normalize(sub(l.p, hit));                                                      // :667
    V3 V = normalize(ray.d);                                                              // :668
    V3 diff = v3(0, 0, 0);
    if (m.is_diffuse()) {
        float angle = dot(n, L);                                                          // :672
        diff = mulf(m.kd, cs_max(0, angle));                                              // :677-678
    }
    V3 spec = v3(0, 0, 0);
    if (m.has_spec()) {
        V3 rv = sub(L, mulf(n, 2 * dot(L, n)));                                           // :683-684
        float s = dot(V, normalize(rv));                                                  // :685-688
        float pw = (float)std::pow((double)cs_max(0, s), (double)m.n);                    // :691
        spec = mulv(m.ks, splat(pw));
        if (cx.count_chain) cx.cnt.shade_specular++;
    }
    return add(diff, spec);                                                               // :694
}

// CalculateReflectionRay :718-720
static V3 reflect_dir(V3 v, V3 n) { return sub(v, mulf(n, 2 * dot(v, n))); }

struct Trace { float d; V3 col; };                                                        // TraceResult :206-231

static V3 trace_secondary(Ctx& cx, V3 origin, V3 dir, int bounce, uint32_t level);

// The part of TraceSphere after the intersection test: :839-875
static Trace shade_sphere(Ctx& cx, const Ray& ray, const Sphere& s, Isect is, int bounce, uint32_t level) {
    const bool on_chain = cx.count_chain;
    const Scene& sc = *cx.sc;
    Trace r; r.d = is.d; r.col = v3(0, 0, 0);
    if (!is.hit || is.d - 0.01f <= 0) return r;                                           // :839-840
    if (bounce > sc.cap) return r;                                                        // :843 (black)
    V3 hit = add(ray.o, mulf(ray.d, is.d));                                               // :846
    V3 col = v3(0, 0, 0);
    if (on_chain) cx.cnt.shaded_hits++;
    if (s.m.is_mirror()) {                                                                // :850
        bounce++;
        V3 rd = reflect_dir(ray.d, normalize(sub(hit, s.c)));                             // :852-855
        if (on_chain) cx.cnt.shade_mirror++;
        col = add(col, mulv(trace_secondary(cx, hit, rd, bounce, level + 1), s.m.km)); // :857
    }
    if (s.m.is_diffuse()) {                                                               // :862
        uint32_t li = 0;
        for (const Light& l : sc.lights) {
            if (on_chain) { cx.cnt.shadow++; cx.cnt.shade_diffuse++; }
            cx.cnt.faithful_shadow++;
            float I = intersect_shadow_light(cx, hit, l);                                 // :864
            if (on_chain) cx.hash += event_hash(level, 3, li, I == 0.0f ? 1u : 0u);
            if (on_chain) log_shadow(cx, level, li, hit, l);
            V3 irgb = splat(I);                                                           // :865
            float att = 1 / is.d * is.d;                                                  // :866 ((1/d)*d)
            V3 n = normalize(sub(hit, s.c));                                              // :706
            col = add(col, mulv(mulf(irgb, att), shape_phong(cx, hit, ray, n, s.m, l)));  // :868-869
            li++;
        }
    }
    col = add(col, mulv(sc.ambient, s.m.ka));                                             // :873
    r.col = col; return r;
}

// The part of TracePlane after the intersection test: :731-779
static Trace shade_plane(Ctx& cx, const Ray& ray, const Plane& p, Isect is, int bounce, uint32_t level) {
    const bool on_chain = cx.count_chain;
    const Scene& sc = *cx.sc;
    Trace r; r.d = is.d; r.col = v3(0, 0, 0);
    if (!is.hit || is.d - 0.01f <= 0) return r;                                           // :731-732
    if (bounce > sc.cap) { r.col = v3(1, 1, 1); return r; }                               // :734 (white)
    V3 hit = add(ray.o, mulf(ray.d, is.d));                                               // :736
    V3 col = v3(0, 0, 0);
    if (on_chain) cx.cnt.shaded_hits++;
    if (p.m.is_mirror()) {                                                                // :739
        bounce++;
        V3 rd = reflect_dir(ray.d, p.n);                                                  // :741-744
        if (on_chain) cx.cnt.shade_mirror++;
        col = add(col, mulv(trace_secondary(cx, hit, rd, bounce, level + 1), p.m.km)); // :746-747
    }
    if (p.m.is_diffuse()) {                                                               // :750
        uint32_t li = 0;
        for (const Light& l : sc.lights) {
            if (on_chain) { cx.cnt.shadow++; cx.cnt.shade_diffuse++; }
            cx.cnt.faithful_shadow++;
            float I = intersect_shadow_light(cx, hit, l);                                 // :752
            if (on_chain) cx.hash += event_hash(level, 3, li, I == 0.0f ? 1u : 0u);
            if (on_chain) log_shadow(cx, level, li, hit, l);
            V3 irgb = splat(I);                                                           // :753
            float att = (float)(1 / std::pow((double)is.d, 2.0));                         // :754
            V3 tile = v3(1, 1, 1);                                                        // :756
            if (p.tiled) {                                                                // :757 (always true, :289)
                V3 e1 = normalize(cross(p.n, v3(1.0f, 0.0f, 0.0f)));                      // :760
                if (e1.x == 0.0f && e1.y == 0.0f && e1.z == 0.0f)                         // :761
                    e1 = normalize(cross(p.n, v3(0, 0, 1)));                              // :762
                V3 e2 = normalize(cross(p.n, e1));                                        // :765
                float u = dot(e1, hit);                                                   // :766
                float v = dot(e2, hit);                                                   // :767
                int32_t cb = (int32_t)(((uint32_t)cs_f2i(u) + (uint32_t)cs_f2i(v)) & 1u); // :769 (+ before &)
                tile = splat((float)cb);                                                  // :770
            }
            V3 ph = shape_phong(cx, hit, ray, p.n, p.m, l);                               // :652-654
            col = add(col, vmaxf(mulv(mulv(mulf(irgb, att), ph), tile), 0.0f));           // :773-775
            li++;
        }
    }
    col = add(col, mulv(sc.ambient, p.m.ka));                                             // :778
    r.col = col; return r;
}

// TraceSecondaryRay :789-826.  faithful: shade every primitive then select (as written).
// nearest-first: select by distance, then shade only the winner (bit-identical, tested).
static V3 trace_secondary(Ctx& cx, V3 origin, V3 dir, int bounce, uint32_t level) {
    const Scene& sc = *cx.sc;
    Ray ray = {origin, dir};                                                              // :793 / :814
    if (cx.count_chain) cx.cnt.secondary++;
    cx.cnt.faithful_secondary++;

    float closest_s = INFINITY; int sel_s = -1; V3 col_s = v3(0, 0, 0); Isect is_s = {false, 0};
    for (int i = 0; i < (int)sc.spheres.size(); i++) {                                    // :792
        Isect is = intersects_sphere(cx, ray, sc.spheres[i], 0.0f);                       // :836
        Trace tr; tr.d = is.d; tr.col = v3(0, 0, 0);
        if (cx.faithful) tr = shade_sphere(cx, ray, sc.spheres[i], is, bounce, level);    // :794-798
        if (tr.d - 0.01f > 0 && tr.d - 0.01f < closest_s) {                               // :804
            closest_s = tr.d; sel_s = i; is_s = is; col_s = tr.col;                       // :805-806
        }
    }
    float closest_p = INFINITY; int sel_p = -1; V3 col_p = v3(0, 0, 0); Isect is_p = {false, 0};
    for (int i = 0; i < (int)sc.planes.size(); i++) {                                     // :812
        Isect is = intersect_plane(cx, ray, sc.planes[i]);                                // :730
        Trace tr; tr.d = is.d; tr.col = v3(0, 0, 0);
        if (cx.faithful) tr = shade_plane(cx, ray, sc.planes[i], is, bounce, level);      // :813-817
        if (tr.d > 0 && tr.d < closest_p) {                                               // :819
            closest_p = tr.d; sel_p = i; is_p = is; col_p = tr.col;                       // :820-821
        }
    }
    bool pick_s = closest_s < closest_p;                                                  // :825
    if (cx.faithful) return pick_s ? col_s : col_p;
    uint32_t code = pick_s ? (uint32_t)sel_s : (sel_p >= 0 ? (uint32_t)(sc.spheres.size() + sel_p) : 0xFFFFFFFFu);
    float dsel = pick_s ? closest_s : (sel_p >= 0 ? closest_p : 0.0f);
    cx.hash += event_hash(level, 2, code, fbits(dsel));
    log_ray(cx, 1, level, 0, (int32_t)code, origin, dir, dsel);
    if (pick_s) return shade_sphere(cx, ray, sc.spheres[sel_s], is_s, bounce, level).col;
    if (sel_p >= 0) return shade_plane(cx, ray, sc.planes[sel_p], is_p, bounce, level).col;
    return v3(0, 0, 0);
}

struct Camera { V3 pos, right, up, fwd, view; };                                          // rt_camera (15 floats)

// Jitter spec (extension, not in the reference). spp == 1 -> no jitter.
static inline uint32_t pcg_hash(uint32_t v) {
    uint32_t s = v * 747796405u + 2891336453u;
    uint32_t w = ((s >> ((s >> 28) + 4)) ^ s) * 277803737u;
    return (w >> 22) ^ w;
}

// ShiftColor :1046-1052
static int32_t shift_color(V3 c) {
    int32_t r = cs_d2i(std::floor((double)(cs_clamp(c.x, 0.0f, 1.0f) * 255.0f)));
    int32_t g = cs_d2i(std::floor((double)(cs_clamp(c.y, 0.0f, 1.0f) * 255.0f)));
    int32_t b = cs_d2i(std::floor((double)(cs_clamp(c.z, 0.0f, 1.0f) * 255.0f)));
    return (int32_t)(((uint32_t)(uint8_t)r << 16) | ((uint32_t)(uint8_t)g << 8) | (uint32_t)(uint8_t)b);
}

// TracePixel :962-1002 for one sample position (fx, fy) (= (x, y) when spp == 1).
static V3 trace_sample(Ctx& cx, const Camera& cam, float fx, float fy, int w, int h, int32_t* aov_id, float* aov_t) {
    const Scene& sc = *cx.sc;
    float u = fx / (float)w - 0.5f;                                                       // :964
    float v = fy / (float)h - 0.5f;
    V3 local = mulv(v3(u, v, 1.0f), cam.view);                                            // :965
    V3 vp = add(add(add(cam.pos, mulf(cam.right, local.x)), mulf(cam.up, local.y)), mulf(cam.fwd, local.z)); // :967-969
    Ray ray = {cam.pos, normalize(sub(vp, cam.pos))};                                     // :971
    cx.cnt.primary++;

    V3 col_s = v3(0, 0, 0); float near_s = INFINITY; int sel_s = -1; Isect is_s = {false, 0};
    for (int i = 0; i < (int)sc.spheres.size(); i++) {                                    // :975
        Isect is = intersects_sphere(cx, ray, sc.spheres[i], 0.0f);
        Trace tr; tr.d = is.d; tr.col = v3(0, 0, 0);
        if (cx.faithful) tr = shade_sphere(cx, ray, sc.spheres[i], is, 0, 0);             // :976
        if (tr.d > 0 && near_s > tr.d) {                                                  // :977
            near_s = tr.d; sel_s = i; is_s = is; col_s = tr.col;                          // :978-979
        }
    }
    V3 col_p = v3(0, 0, 0); float near_p = INFINITY; int sel_p = -1; Isect is_p = {false, 0};
    for (int i = 0; i < (int)sc.planes.size(); i++) {                                     // :985
        Isect is = intersect_plane(cx, ray, sc.planes[i]);
        Trace tr; tr.d = is.d; tr.col = v3(0, 0, 0);
        if (cx.faithful) tr = shade_plane(cx, ray, sc.planes[i], is, 0, 0);               // :986
        if (tr.d > 0 && near_p > tr.d) {                                                  // :987
            near_p = tr.d; sel_p = i; is_p = is; col_p = tr.col;                          // :988-989
        }
    }
    bool pick_s = near_s < near_p;                                                        // :993
    uint32_t code = pick_s ? (uint32_t)sel_s : (sel_p >= 0 ? (uint32_t)(sc.spheres.size() + sel_p) : 0xFFFFFFFFu);
    float dsel = pick_s ? near_s : (sel_p >= 0 ? near_p : 0.0f);
    if (aov_id) *aov_id = (int32_t)code;
    if (aov_t) *aov_t = dsel;
    if (cx.faithful) return pick_s ? col_s : col_p;
    cx.hash += event_hash(0, 1, code, fbits(dsel));
    log_ray(cx, 0, 0, 0, (int32_t)code, ray.o, ray.d, dsel);
    if (pick_s) return shade_sphere(cx, ray, sc.spheres[sel_s], is_s, 0, 0).col;
    if (sel_p >= 0) return shade_plane(cx, ray, sc.planes[sel_p], is_p, 0, 0).col;
    return v3(0, 0, 0);
}

static void render_pixel(Ctx& cx, const Camera& cam, int x, int y, int w, int h, int spp, uint32_t seed,
                         int32_t* out_px, uint32_t* out_hash, int32_t* aov_id, float* aov_t) {
    cx.hash = 0;
    V3 col;
    if (spp <= 1) {
        col = trace_sample(cx, cam, (float)x, (float)y, w, h, aov_id, aov_t);
    } else {
        V3 acc = v3(0, 0, 0);
        for (int s = 0; s < spp; s++) {
            uint32_t k = ((uint32_t)y * (uint32_t)w + (uint32_t)x) * (uint32_t)spp + (uint32_t)s;
            uint32_t h1 = pcg_hash(k ^ seed), h2 = pcg_hash(h1);
            float jx = (float)(h1 >> 8) * 5.9604644775390625e-08f;   // 2^-24
            float jy = (float)(h2 >> 8) * 5.9604644775390625e-08f;
            V3 c = trace_sample(cx, cam, (float)x + jx, (float)y + jy, w, h, s == 0 ? aov_id : nullptr, s == 0 ? aov_t : nullptr);
            acc = add(acc, c);
        }
        col = mulf(acc, 1.0f / (float)spp);
    }
    *out_px = shift_color(col);                                                           // :1000 -> :1038
    if (out_hash) *out_hash = cx.hash;
}

static Scene build_scene(const float* spheres, int ns, const float* planes, int np, const float* lights, int nl,
                         const float* ambient, int cap) {
    Scene sc;
    sc.spheres.resize(ns); sc.planes.resize(np); sc.lights.resize(nl);
    for (int i = 0; i < ns; i++) {
        const float* f = spheres + 18 * i;
        sc.spheres[i].c = v3(f[0], f[1], f[2]); sc.spheres[i].r = f[3]; sc.spheres[i].m = load_mat(f + 4); sc.spheres[i].r2 = f[17];
    }
    for (int i = 0; i < np; i++) {
        const float* f = planes + 20 * i;
        sc.planes[i].c = v3(f[0], f[1], f[2]); sc.planes[i].n = v3(f[3], f[4], f[5]); sc.planes[i].m = load_mat(f + 6);
        sc.planes[i].tiled = true;                                                        // :289 — isTiled is ALWAYS true
    }
    for (int i = 0; i < nl; i++) { const float* f = lights + 4 * i; sc.lights[i].p = v3(f[0], f[1], f[2]); sc.lights[i].i = f[3]; }
    sc.ambient = v3(ambient[0], ambient[1], ambient[2]);
    sc.cap = cap;
    return sc;
}

static Camera load_cam(const float* c) {
    Camera cam;
    cam.pos = v3(c[0], c[1], c[2]); cam.right = v3(c[3], c[4], c[5]); cam.up = v3(c[6], c[7], c[8]);
    cam.fwd = v3(c[9], c[10], c[11]); cam.view = v3(c[12], c[13], c[14]);
    return cam;
}

}  // namespace

extern "C" {

// mode: 0 = faithful (shade every intersected primitive then select; per-column fork/join like Tick :898-901),
//       1 = nearest-first (select then shade; OpenMP over rows).
// subset: optional list of pixel indices (y*w+x) to render; pixels/hash/aov arrays are then n_subset long.
// counters: N_COUNTERS uint64 (see struct Counters order), nullable. hash/aov_id/aov_t nullable.
int orc_render(const float* spheres, int ns, const float* planes, int np, const float* lights, int nl,
               const float* ambient, const float* cam15, int w, int h, int max_depth, int spp, uint32_t seed,
               int mode, int threads, int32_t* pixels, uint64_t* counters, uint32_t* hash,
               int32_t* aov_id, float* aov_t, const int32_t* subset, int n_subset) {
    if (w <= 0 || h <= 0 || !pixels || ns < 0 || np < 0 || nl < 0) return -1;
    Scene sc = build_scene(spheres, ns, planes, np, lights, nl, ambient, max_depth);
    Camera cam = load_cam(cam15);
    bool faithful = (mode == 0);
#ifdef _OPENMP
    int nth = threads > 0 ? threads : omp_get_max_threads();
#else
    int nth = 1;
#endif
    std::vector<Counters> tc((size_t)nth);
    if (subset) {
#pragma omp parallel num_threads(nth)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            Ctx cx; cx.sc = &sc; cx.hash = 0; cx.faithful = faithful; cx.count_chain = !faithful;
#pragma omp for schedule(dynamic, 64)
            for (int i = 0; i < n_subset; i++) {
                int idx = subset[i]; int x = idx % w, y = idx / w;
                render_pixel(cx, cam, x, y, w, h, spp, seed, pixels + i, hash ? hash + i : nullptr,
                             aov_id ? aov_id + i : nullptr, aov_t ? aov_t + i : nullptr);
            }
            tc[tid].add(cx.cnt);
        }
    } else if (faithful) {
        // Tick :898-901 — serial loop over columns, Parallel.For over the rows of each column.
        std::memset(pixels, 0, (size_t)w * h * 4);                                        // screen.Clear(0) :890
#pragma omp parallel num_threads(nth)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            Ctx cx; cx.sc = &sc; cx.hash = 0; cx.faithful = true; cx.count_chain = false;
            for (int x = 0; x < w; x++) {
#pragma omp for schedule(dynamic, 16)
                for (int y = 0; y < h; y++) {
                    size_t i = (size_t)y * w + x;
                    render_pixel(cx, cam, x, y, w, h, spp, seed, pixels + i, hash ? hash + i : nullptr,
                                 aov_id ? aov_id + i : nullptr, aov_t ? aov_t + i : nullptr);
                }   // implicit barrier = the per-column join
            }
            tc[tid].add(cx.cnt);
        }
    } else {
#pragma omp parallel num_threads(nth)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            Ctx cx; cx.sc = &sc; cx.hash = 0; cx.faithful = false; cx.count_chain = true;
#pragma omp for schedule(dynamic, 4)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    size_t i = (size_t)y * w + x;
                    render_pixel(cx, cam, x, y, w, h, spp, seed, pixels + i, hash ? hash + i : nullptr,
                                 aov_id ? aov_id + i : nullptr, aov_t ? aov_t + i : nullptr);
                }
            tc[tid].add(cx.cnt);
        }
    }
    if (counters) {
        Counters tot; for (auto& c : tc) tot.add(c);
        uint64_t v[N_COUNTERS] = {tot.primary, tot.shadow, tot.secondary, tot.sphere_tests, tot.sphere_disc_pos,
                                  tot.plane_tests, tot.shade_diffuse, tot.shade_specular, tot.shade_mirror,
                                  tot.shaded_hits, tot.faithful_secondary, tot.faithful_shadow};
        std::memcpy(counters, v, sizeof(v));
    }
    return 0;
}

// Brute-force single-ray queries used by the LBVH-equality harness.
// kind 0: primary fold (:975-981 strict '<', d>0); kind 1: secondary fold (:792-808, offset compare, order
// dependent); kind 2: shadow any-hit with eps=0.001 (:573-582). Returns id (-1 none) and t per ray.
int orc_query_spheres(const float* spheres, int ns, const float* rays6, int n_rays, int kind,
                      int32_t* out_id, float* out_t) {
    const float zero3[3] = {0, 0, 0};
    Scene sc = build_scene(spheres, ns, nullptr, 0, nullptr, 0, zero3, 0);
#pragma omp parallel for schedule(dynamic, 256)
    for (int r = 0; r < n_rays; r++) {
        Ctx cx; cx.sc = &sc; cx.hash = 0; cx.faithful = false; cx.count_chain = false;
        Ray ray = {v3(rays6[6 * r], rays6[6 * r + 1], rays6[6 * r + 2]), v3(rays6[6 * r + 3], rays6[6 * r + 4], rays6[6 * r + 5])};
        int sel = -1; float best = INFINITY;
        if (kind == 0) {
            for (int i = 0; i < ns; i++) { Isect is = intersects_sphere(cx, ray, sc.spheres[i], 0.0f); if (is.d > 0 && best > is.d) { best = is.d; sel = i; } }
        } else if (kind == 1) {
            for (int i = 0; i < ns; i++) { Isect is = intersects_sphere(cx, ray, sc.spheres[i], 0.0f); if (is.d - 0.01f > 0 && is.d - 0.01f < best) { best = is.d; sel = i; } }
        } else {
            for (int i = 0; i < ns; i++) if (intersects_sphere(cx, ray, sc.spheres[i], 0.001f).hit) { sel = 1; }
            best = 0.0f; if (sel < 0) sel = 0;   // id: 1 = occluded, 0 = clear
        }
        out_id[r] = sel; out_t[r] = (sel >= 0 && kind != 2) ? best : 0.0f;
    }
    return 0;
}

// Ray log of the listed pixels (nearest-first mode, spp = 1): the checker of rt_ray_log. Records grouped by pixel in list
// order; within a pixel in creation order of the recursion. Returns the number of records; writes at most max_records.
int orc_ray_log(const float* spheres, int ns, const float* planes, int np, const float* lights, int nl,
                const float* ambient, const float* cam15, int w, int h, int max_depth,
                const uint32_t* pixels, int n_pixels, void* out_records, int max_records) {
    if (w <= 0 || h <= 0 || n_pixels < 0 || (n_pixels && !pixels)) return -1;
    Scene sc = build_scene(spheres, ns, planes, np, lights, nl, ambient, max_depth);
    Camera cam = load_cam(cam15);
    std::vector<RayRecord> log;
    Ctx cx; cx.sc = &sc; cx.hash = 0; cx.faithful = false; cx.count_chain = true; cx.log = &log;
    for (int i = 0; i < n_pixels; i++) {
        uint32_t p = pixels[i];
        if (p >= (uint32_t)w * (uint32_t)h) return -1;
        cx.log_pixel = p;
        trace_sample(cx, cam, (float)(p % (uint32_t)w), (float)(p / (uint32_t)w), w, h, nullptr, nullptr);
    }
    size_t take = log.size() < (size_t)max_records ? log.size() : (size_t)max_records;
    if (take && out_records) std::memcpy(out_records, log.data(), take * sizeof(RayRecord));
    return (int)log.size();
}

int orc_pack_color(float r, float g, float b) { return shift_color(v3(r, g, b)); }
int orc_num_counters(void) { return N_COUNTERS; }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
