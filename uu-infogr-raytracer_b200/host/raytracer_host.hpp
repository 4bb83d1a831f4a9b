// raytracer_host.hpp — C++ mirror of the reference's host classes above the C ABI (include/rtb200.h).
//
// The reference host is C# (Raytracer/RayTracer.cs, surface.cs); no .NET toolchain exists in the build image, so the
// host side that can be compiled and tested here is this C++ restatement of the same classes with the same members:
//   Surface   (surface.cs:7-20, 43-46)   width, height, pixels[], Clear
//   RayTracer (RayTracer.cs:437-1062)    scene arrays :441-469, camera state :494-523, OnKeyPress :543-554,
//                                        OnMouseMove :1058-1061, Tick :886-901
// with the private trace methods (:573-876, :962-1052) replaced by rt_render().  The C# binding a maintainer would use
// is csharp/NativeRayTracer.cs; INTEGRATION.md shows the patch.  Host arithmetic follows the C# float/double mix exactly.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtb200.h"

namespace rthost {

struct Vector3 { float X, Y, Z; };
inline Vector3 operator+(Vector3 a, Vector3 b) { return {a.X + b.X, a.Y + b.Y, a.Z + b.Z}; }
inline Vector3 operator-(Vector3 a, Vector3 b) { return {a.X - b.X, a.Y - b.Y, a.Z - b.Z}; }
inline Vector3 operator*(Vector3 a, Vector3 b) { return {a.X * b.X, a.Y * b.Y, a.Z * b.Z}; }
inline Vector3 Cross(Vector3 l, Vector3 r) {           // OpenTK Vector3.Cross
    return {(l.Y * r.Z) - (l.Z * r.Y), (l.Z * r.X) - (l.X * r.Z), (l.X * r.Y) - (l.Y * r.X)};
}

struct Material {                                      // RayTracer.cs:60-159
    Vector3 diffuseColor, ambientColor, specularColor; float specularity; Vector3 mirrorColor;
    static Material Diffuse(Vector3 c) { return {c, c, {0, 0, 0}, 0.0f, {0, 0, 0}}; }                           // :117
    static Material Plastic(Vector3 c, float n = 1.0f) { return {c, c, {0.4f, 0.4f, 0.4f}, n, {0, 0, 0}}; }    // :127
    static Material Metal(Vector3 c, float n = 1.0f) { return {c, c, c, n, {0, 0, 0}}; }                       // :137
    static Material Mirror(Vector3 m) { return {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 0.0f, m}; }                   // :146
    static Material DiffuseMirror(Vector3 c, Vector3 m) { return {c, c, {0, 0, 0}, 0.0f, m}; }                 // :156
};
struct Sphere {                                        // :308-338 — 18 floats, the layout rt_set_scene expects
    Vector3 center; float radius; Material material; float radiusSquared;
    Sphere(Vector3 c, float r, Material m) : center(c), radius(r), material(m), radiusSquared(r * r) {}
};
struct Plane {                                         // :260-303
    Vector3 center, normal; Material material; bool isTiled;
    Plane(Vector3 c, Vector3 n, Material m, bool = false) : center(c), normal(n), material(m), isTiled(true) {}  // :289
};
struct Light { Vector3 position; float intensity; };   // :236-255
static_assert(sizeof(Sphere) == 18 * sizeof(float), "Sphere must be 18 packed floats");
static_assert(sizeof(Light) == 4 * sizeof(float), "Light must be 4 packed floats");

class Surface {                                        // surface.cs:7-20
public:
    int width, height;
    std::vector<int32_t> pixels;
    Surface(int w, int h) : width(w), height(h), pixels((size_t)w * h, 0) {}
    void Clear(int c) { for (auto& p : pixels) p = c; }                                                          // :43-46
};

enum class Keys { W, A, S, D, Space, LeftShift, RightShift, Other };

class RayTracer {
public:
    std::vector<Sphere> _spheres = {                                                                             // :441-445
        Sphere({2.5f, 0, 8}, 1.0f, Material::Diffuse({1.0f, 0, 0})),
        Sphere({3, 0, 5}, 1.0f, Material::Plastic({0, 1, 0})),
        Sphere({-3, 1, 8}, 1.0f, Material::Mirror({1, 1, 1})),
    };
    std::vector<Light> _lights = {{{-3, 1, -3}, 1.0f}, {{33, 1, 10}, 1.0f}};                                    // :450-453
    std::vector<Plane> _planes = {Plane({0, -1.0f, 0}, {0, 1, 0},                                                // :458-465
                                        Material{{1, 1, 1}, {0.5f, 0.5f, 0.5f}, {1, 1, 1}, 0.5f, {1, 1, 1}}, true)};
    Vector3 _ambientLightColor = {43.0f / 255.0f, 43.0f / 255.0f, 43.0f / 255.0f};                              // :469
    static constexpr float NearClip = 0.3f;                                                                      // :481
    static constexpr float FieldOfView = 60.0f;                                                                  // :486
    static constexpr int ReflectionRecursionLimit = 32;                                                          // :490
    Vector3 _cameraPosition = {0.0f, 0.0f, 0.0f};                                                               // :494
    float _yaw = 0.0f, _pitch = 0.0f;                                                                            // :498-502
    Surface& screen;                                                                                             // :506
    rt_stats last_stats{};

    Vector3 CameraForwardDirection() const {                                                                     // :511-513
        return {(float)(std::cos((double)_pitch) * std::sin((double)_yaw)), (float)-std::sin((double)_pitch),
                (float)(std::cos((double)_pitch) * std::cos((double)_yaw))};
    }
    Vector3 CameraRightDirection() const { return {(float)std::cos((double)_yaw), 0, (float)-std::sin((double)_yaw)}; }   // :517-518
    Vector3 CameraUpDirection() const { return Cross(CameraRightDirection(), CameraForwardDirection()); }       // :522-523

    explicit RayTracer(Surface& s, int n_devices = 1) : screen(s) {                                              // :535-537
        int rc = rt_create(&_ctx, nullptr, n_devices);
        if (rc != RT_OK) throw std::runtime_error(std::string("rt_create: ") + rt_last_error(nullptr));
        rt_host_register(_ctx, screen.pixels.data(), screen.pixels.size() * sizeof(int32_t));   // optional; ignore failure
        UploadScene();
    }
    ~RayTracer() {
        if (_ctx) { rt_host_unregister(_ctx, screen.pixels.data()); rt_destroy(_ctx); }
    }
    RayTracer(const RayTracer&) = delete;
    RayTracer& operator=(const RayTracer&) = delete;

    void UploadScene(int accel = RT_ACCEL_AUTO) {
        std::vector<float> p(_planes.size() * 20);
        for (size_t i = 0; i < _planes.size(); i++) {
            const Plane& pl = _planes[i]; const Material& m = pl.material; float* o = &p[i * 20];
            o[0] = pl.center.X; o[1] = pl.center.Y; o[2] = pl.center.Z; o[3] = pl.normal.X; o[4] = pl.normal.Y; o[5] = pl.normal.Z;
            o[6] = m.diffuseColor.X; o[7] = m.diffuseColor.Y; o[8] = m.diffuseColor.Z;
            o[9] = m.ambientColor.X; o[10] = m.ambientColor.Y; o[11] = m.ambientColor.Z;
            o[12] = m.specularColor.X; o[13] = m.specularColor.Y; o[14] = m.specularColor.Z; o[15] = m.specularity;
            o[16] = m.mirrorColor.X; o[17] = m.mirrorColor.Y; o[18] = m.mirrorColor.Z; o[19] = pl.isTiled ? 1.0f : 0.0f;
        }
        const float amb[3] = {_ambientLightColor.X, _ambientLightColor.Y, _ambientLightColor.Z};
        Check(rt_set_scene(_ctx, reinterpret_cast<const float*>(_spheres.data()), (int)_spheres.size(), p.data(), (int)_planes.size(),
                           reinterpret_cast<const float*>(_lights.data()), (int)_lights.size(), amb, accel));
    }

    void OnKeyPress(Keys key) {                                                                                  // :543-554
        const Vector3 moveScaler = {0.05f, 0.05f, 0.05f};
        switch (key) {
            case Keys::W: _cameraPosition = _cameraPosition + CameraForwardDirection() * moveScaler; break;
            case Keys::A: _cameraPosition = _cameraPosition - CameraRightDirection() * moveScaler; break;
            case Keys::S: _cameraPosition = _cameraPosition - CameraForwardDirection() * moveScaler; break;
            case Keys::D: _cameraPosition = _cameraPosition + CameraRightDirection() * moveScaler; break;
            case Keys::Space: _cameraPosition = _cameraPosition - CameraUpDirection() * moveScaler; break;
            case Keys::LeftShift: case Keys::RightShift: _cameraPosition = _cameraPosition + CameraUpDirection() * moveScaler; break;
            default: break;
        }
    }
    void OnMouseMove(float deltaX, float deltaY) { _yaw += deltaX / 360; _pitch += deltaY / 360; }               // :1058-1061

    rt_camera FrameCamera() const {                                                                              // :892-896
        const float degToRad = 3.14159274f / 180.0f;                                  // MathHelper.DegreesToRadians (OpenTK)
        float planeHeight = NearClip * (float)std::tan((double)(FieldOfView * 0.5f * degToRad)) * 2;               // :892
        float aspectRatio = (float)screen.width / screen.height;                                                  // :893
        float planeWidth = planeHeight * aspectRatio;                                                             // :894
        rt_camera cam;
        Put(cam.pos, _cameraPosition); Put(cam.right, CameraRightDirection()); Put(cam.up, CameraUpDirection());
        Put(cam.forward, CameraForwardDirection());
        cam.view_params[0] = planeWidth; cam.view_params[1] = planeHeight; cam.view_params[2] = NearClip;         // :896
        return cam;
    }

    void Tick() {                                                                                                // :886-901
        // screen.Clear(0) (:890) is subsumed: the backend writes every pixel.
        rt_camera cam = FrameCamera();
        // for (x) Parallel.For(y => TracePixel(x, y, viewParams))  (:898-901)  ==>
        Check(rt_render(_ctx, &cam, screen.width, screen.height, ReflectionRecursionLimit, 1, 0u, screen.pixels.data(), &last_stats));
    }

    // The data of the DEBUG_ENABLE overlay: one TracedRay (:424-435) per ray of the listed pixels' chains (:914-933 draws a sample).
    std::vector<rt_ray_record> TracedRays(const std::vector<uint32_t>& pixelIndices) {
        rt_camera cam = FrameCamera();
        int n = 0;
        Check(rt_ray_log(_ctx, &cam, screen.width, screen.height, ReflectionRecursionLimit, pixelIndices.data(), (int)pixelIndices.size(),
                         nullptr, 0, &n));
        std::vector<rt_ray_record> recs((size_t)n);
        if (n) Check(rt_ray_log(_ctx, &cam, screen.width, screen.height, ReflectionRecursionLimit, pixelIndices.data(),
                                (int)pixelIndices.size(), recs.data(), n, &n));
        return recs;
    }

private:
    rt_context* _ctx = nullptr;
    static void Put(float* dst, Vector3 v) { dst[0] = v.X; dst[1] = v.Y; dst[2] = v.Z; }
    void Check(int rc) const {
        if (rc != RT_OK) throw std::runtime_error(std::string("rtb200 error ") + std::to_string(rc) + ": " + rt_last_error(_ctx));
    }
};

}  // namespace rthost
