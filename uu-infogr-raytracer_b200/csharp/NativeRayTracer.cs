// NativeRayTracer.cs — P/Invoke binding of librtb200.so for the reference host (Raytracer/, net6.0).
//
// Drop this file into Raytracer/ next to RayTracer.cs, ship librtb200.so beside the executable, and apply the
// ~10-line patch shown in INTEGRATION.md to RayTracer.cs (constructor: create + upload scene; Tick(): replace the
// pixel loop :898-901 by Render()).  Everything else — scene arrays (:441-465), camera state and input handlers
// (:494-554, :1058-1061), Surface (surface.cs) and the OpenTK display path (template.cs) — stays as it is.
//
// This file cannot be compiled or run in the build image (no .NET toolchain there); it is kept deliberately thin so
// that review == verification: every call maps 1:1 to a function of include/rtb200.h.
using System;
using System.Runtime.InteropServices;
using OpenTK.Mathematics;

namespace Template;

internal sealed unsafe class NativeRayTracer : IDisposable {
    private const string Lib = "rtb200";    // librtb200.so / rtb200.dll

    [StructLayout(LayoutKind.Sequential)]
    private struct RtCamera {                // include/rtb200.h: rt_camera (15 floats)
        public Vector3 pos, right, up, forward, viewParams;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct RtStats {                  // include/rtb200.h: rt_stats
        public ulong primary, shadow, secondary;
        public float kernelMs, gatherMs, d2hMs;
    }

    [DllImport(Lib)] private static extern int rt_create(out IntPtr ctx, int[]? deviceIds, int nDevices);
    [DllImport(Lib)] private static extern int rt_set_scene(IntPtr ctx, float* spheres, int nSpheres, float* planes, int nPlanes,
                                                            float* lights, int nLights, float* ambient, int accel);
    [DllImport(Lib)] private static extern int rt_render(IntPtr ctx, ref RtCamera cam, int width, int height, int maxDepth, int spp,
                                                         uint seed, int* hostPixels, out RtStats stats);
    [StructLayout(LayoutKind.Sequential)]
    public struct RtRayRecord {              // include/rtb200.h: rt_ray_record (64 bytes) — one TracedRay (RayTracer.cs:424-435) per ray
        public Vector3 origin, direction, hitPoint;
        public float distance;
        public int hit;                      // sphere index, spheres.Length + plane index, -1 none (shadow: nearest occluder)
        public uint kind;                    // (uint)RayKind: 0 Primary, 1 Secondary, 2 Shadow (:343-362)
        public uint pixel, level, light, reserved;
    }

    [DllImport(Lib)] private static extern int rt_ray_log(IntPtr ctx, ref RtCamera cam, int width, int height, int maxDepth, uint* pixels,
                                                          int nPixels, RtRayRecord* records, int maxRecords, out int nRecords);
    [DllImport(Lib)] private static extern int rt_selftest(IntPtr ctx, int test, out ulong nChecked, out ulong nMismatch);
    [DllImport(Lib)] private static extern int rt_set_option(IntPtr ctx, int option, int value);
    [DllImport(Lib)] private static extern int rt_get_info(IntPtr ctx, int what, out ulong value);
    [DllImport(Lib)] private static extern int rt_update_spheres(IntPtr ctx, float* spheres, int first, int count);
    [DllImport(Lib)] private static extern int rt_render_mapped(IntPtr ctx, ref RtCamera cam, int width, int height, int maxDepth, int spp,
                                                                uint seed, void* mappedDevPixels, ulong mappedBytes, out RtStats stats);
    [DllImport(Lib)] private static extern int rt_gl_register_buffer(IntPtr ctx, uint glBuffer, out IntPtr resource);
    [DllImport(Lib)] private static extern int rt_gl_unregister_buffer(IntPtr ctx, IntPtr resource);
    [DllImport(Lib)] private static extern int rt_render_gl(IntPtr ctx, ref RtCamera cam, int width, int height, int maxDepth, int spp,
                                                            uint seed, IntPtr resource, out RtStats stats);
    [DllImport(Lib)] private static extern int rt_host_register(IntPtr ctx, void* hostPtr, ulong bytes);
    [DllImport(Lib)] private static extern int rt_host_unregister(IntPtr ctx, void* hostPtr);
    [DllImport(Lib)] private static extern int rt_destroy(IntPtr ctx);
    [DllImport(Lib)] private static extern IntPtr rt_last_error(IntPtr ctx);

    private IntPtr _ctx;
    private GCHandle _pixelsHandle;          // Surface.pixels pinned for the lifetime of the renderer (zero-copy target)
    private int* _pixels;

    private void Check(int rc) {
        if (rc != 0) throw new InvalidOperationException($"rtb200 error {rc}: {Marshal.PtrToStringAnsi(rt_last_error(_ctx))}");
    }

    /// <summary>Creates the backend on <paramref name="nDevices"/> GPUs (1, 2, 4 or 8) and page-locks the surface.</summary>
    public NativeRayTracer(Surface screen, int nDevices = 1) {
        int rc = rt_create(out _ctx, null, nDevices);
        if (rc != 0) throw new InvalidOperationException($"rtb200 error {rc}: {Marshal.PtrToStringAnsi(rt_last_error(IntPtr.Zero))}");
        _pixelsHandle = GCHandle.Alloc(screen.pixels, GCHandleType.Pinned);
        _pixels = (int*)_pixelsHandle.AddrOfPinnedObject();
        Check(rt_host_register(_ctx, _pixels, (ulong)screen.pixels.Length * sizeof(int)));
    }

    /// <summary>Uploads the scene arrays of RayTracer.cs:441-469. Sphere (18 floats) and Light (4 floats) are blittable and
    /// passed as they lie in memory; Plane holds a bool, so it is packed into 20 floats here.</summary>
    public void SetScene(Sphere[] spheres, Plane[] planes, Light[] lights, Vector3 ambient) {
        float[] p = new float[planes.Length * 20];
        for (int i = 0; i < planes.Length; i++) {
            Plane pl = planes[i]; Material m = pl.material; int o = i * 20;
            p[o + 0] = pl.center.X; p[o + 1] = pl.center.Y; p[o + 2] = pl.center.Z;
            p[o + 3] = pl.normal.X; p[o + 4] = pl.normal.Y; p[o + 5] = pl.normal.Z;
            p[o + 6] = m.diffuseColor.X; p[o + 7] = m.diffuseColor.Y; p[o + 8] = m.diffuseColor.Z;
            p[o + 9] = m.ambientColor.X; p[o + 10] = m.ambientColor.Y; p[o + 11] = m.ambientColor.Z;
            p[o + 12] = m.specularColor.X; p[o + 13] = m.specularColor.Y; p[o + 14] = m.specularColor.Z;
            p[o + 15] = m.specularity;
            p[o + 16] = m.mirrorColor.X; p[o + 17] = m.mirrorColor.Y; p[o + 18] = m.mirrorColor.Z;
            p[o + 19] = pl.isTiled ? 1f : 0f;
        }
        // Sphere = {Vector3 center; float radius; Material (13 floats); float radiusSquared} = 18 sequential floats;
        // Light = {Vector3 position; float intensity}. Add [StructLayout(LayoutKind.Sequential)] to both (and to Material).
        fixed (Sphere* s = spheres) fixed (float* pp = p) fixed (Light* l = lights) {
            float* amb = stackalloc float[3] { ambient.X, ambient.Y, ambient.Z };
            Check(rt_set_scene(_ctx, (float*)s, spheres.Length, pp, planes.Length, (float*)l, lights.Length, amb, 0 /* RT_ACCEL_AUTO */));
        }
    }

    /// <summary>One frame into Surface.pixels — the replacement of the loop RayTracer.cs:898-901.</summary>
    public RtStats Render(Vector3 position, Vector3 right, Vector3 up, Vector3 forward, Vector3 viewParams,
                          int width, int height, int maxDepth) {
        RtCamera cam = new() { pos = position, right = right, up = up, forward = forward, viewParams = viewParams };
        Check(rt_render(_ctx, ref cam, width, height, maxDepth, 1, 0u, _pixels, out RtStats stats));
        return stats;
    }

    /// <summary>Ray records of the listed pixels (y * width + x) — the data the DEBUG_ENABLE overlay draws (RayTracer.cs:914-933):
    /// pick ~500 random pixels outside the debug view, then draw origin -> hitPoint per record, coloured by kind.</summary>
    public RtRayRecord[] RayLog(Vector3 position, Vector3 right, Vector3 up, Vector3 forward, Vector3 viewParams,
                                int width, int height, int maxDepth, uint[] pixels) {
        RtCamera cam = new() { pos = position, right = right, up = up, forward = forward, viewParams = viewParams };
        fixed (uint* px = pixels) {
            Check(rt_ray_log(_ctx, ref cam, width, height, maxDepth, px, pixels.Length, null, 0, out int n));
            RtRayRecord[] records = new RtRayRecord[n];
            fixed (RtRayRecord* r = records)
                Check(rt_ray_log(_ctx, ref cam, width, height, maxDepth, px, pixels.Length, r, n, out n));
            return records;
        }
    }

    /// <summary>Exhaustive on-device check of the library's hand-scheduled fp32 sequences (rt_selftest; 0 = 1/sqrt, 1 = pixel
    /// division). Returns the number of mismatches, which must be 0. Worth running once on new hardware / drivers.</summary>
    public ulong SelfTest(int test) {
        Check(rt_selftest(_ctx, test, out _, out ulong bad));
        return bad;
    }

    /// <summary>rt_set_option (RT_OPT_* of include/rtb200.h), e.g. 7 RT_OPT_HOST_PRECLEARED = 1 when Tick() keeps its screen.Clear(0)
    /// (RayTracer.cs:890): the library then skips its own zero fill of the pixels the sparse return does not copy.</summary>
    public void SetOption(int option, int value) => Check(rt_set_option(_ctx, option, value));

    /// <summary>rt_get_info (RT_INFO_* of include/rtb200.h), e.g. 3 = bytes the last Render copied device -> host.</summary>
    public ulong GetInfo(int what) { Check(rt_get_info(_ctx, what, out ulong v)); return v; }

    /// <summary>Scene mutation (SURVEY §8(f)): re-uploads spheres[first .. first + count), refits the LBVH and rebuilds the shadow bins
    /// on the GPU. The reference's arrays are readonly (RayTracer.cs:441-465); a host that animates them calls this per frame.</summary>
    public void UpdateSpheres(Sphere[] spheres, int first, int count) {
        fixed (Sphere* s = spheres) Check(rt_update_spheres(_ctx, (float*)(s + first), first, count));
    }

    // ---- zero-copy display (INTEGRATION.md §6): the frame goes straight into a GL pixel-unpack buffer, never through Surface.pixels ----
    /// <summary>Registers a GL PixelUnpackBuffer of 4*w*h bytes (render thread, GL context current). Re-register after a resize.</summary>
    public IntPtr GlRegisterBuffer(int glBuffer) { Check(rt_gl_register_buffer(_ctx, (uint)glBuffer, out IntPtr res)); return res; }
    public void GlUnregisterBuffer(IntPtr resource) => Check(rt_gl_unregister_buffer(_ctx, resource));
    /// <summary>map -> render -> unmap; afterwards GL.TexSubImage2D(..., IntPtr.Zero) from the bound buffer replaces template.cs:188-193.</summary>
    public RtStats RenderGl(Vector3 position, Vector3 right, Vector3 up, Vector3 forward, Vector3 viewParams,
                            int width, int height, int maxDepth, IntPtr resource) {
        RtCamera cam = new() { pos = position, right = right, up = up, forward = forward, viewParams = viewParams };
        Check(rt_render_gl(_ctx, ref cam, width, height, maxDepth, 1, 0u, resource, out RtStats stats));
        return stats;
    }
    /// <summary>For hosts that map the buffer themselves: devPixels = cudaGraphicsResourceGetMappedPointer of the mapped PBO.</summary>
    public RtStats RenderMapped(Vector3 position, Vector3 right, Vector3 up, Vector3 forward, Vector3 viewParams,
                                int width, int height, int maxDepth, IntPtr devPixels) {
        RtCamera cam = new() { pos = position, right = right, up = up, forward = forward, viewParams = viewParams };
        Check(rt_render_mapped(_ctx, ref cam, width, height, maxDepth, 1, 0u, (void*)devPixels, 4ul * (ulong)width * (ulong)height, out RtStats stats));
        return stats;
    }

    public void Dispose() {
        if (_ctx == IntPtr.Zero) return;
        rt_host_unregister(_ctx, _pixels);
        rt_destroy(_ctx);
        _ctx = IntPtr.Zero;
        if (_pixelsHandle.IsAllocated) _pixelsHandle.Free();
    }
}
