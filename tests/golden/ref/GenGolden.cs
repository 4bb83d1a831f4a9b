// GenGolden.cs — reference-side fixture generator (VERDICT r01 item 2, SURVEY §8(c)).
//
// Drop this ONE file into the UNMODIFIED reference project directory `Raytracer/` (next to RayTracer.cs) and run, on a machine
// with the .NET 6 SDK (no GPU, no window, no OpenGL context is needed — Surface(w, h) and RayTracer(screen) only allocate):
//
//     cd Raytracer
//     dotnet run -c Release -p:StartupObject=Template.GenGolden -- /path/to/repo/tests/golden/ref
//
// (`-p:StartupObject=...` is needed because template.cs already has a Main.) It constructs the reference's own `RayTracer` on a
// `Surface`, sets the private camera fields `_cameraPosition` / `_yaw` / `_pitch` (RayTracer.cs:494-502) by reflection for each of
// the cameras below, calls the reference's own `Tick()` (RayTracer.cs:886-901) and dumps `screen.pixels` (surface.cs:9-20) as raw
// little-endian int32 — the frames the UNMODIFIED reference renders. It also dumps a probe of the third-party arithmetic on the
// path (OpenTK 4.7.1 Vector3.Normalize / Dot / Cross, System.Math.Pow / Max, (int)float — SURVEY §8 row a19), so that the
// restatement in oracle/rt_oracle.cpp and csrc/rt_math.cuh can be pinned operation by operation.
//
// Outputs (consumed by tests/test_reference_fixtures.py):
//     <name>.bin        w*h int32, row-major, 0x00RRGGBB          one per case of Cases below
//     ref_math.bin      float32 records, layout in MathProbe()
//     manifest.json     the cases (name, w, h, pos, yaw, pitch), the runtime and OpenTK versions
//
// The scene is the reference's hard-coded one (RayTracer.cs:441-469); the recursion cap is its private const 32 (:490).
// Nothing here is part of the product; it is test infrastructure for the oracle.
using System.Reflection;
using System.Runtime.InteropServices;
using System.Text;
using OpenTK.Mathematics;

namespace Template;

internal static class GenGolden
{
    // name, width, height, position, yaw, pitch — keep in sync with CASES in tests/test_reference_fixtures.py (the test checks it)
    private static readonly (string name, int w, int h, float px, float py, float pz, float yaw, float pitch)[] Cases =
    {
        ("ref_default_160x90", 160, 90, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f),
        ("ref_default_192x108", 192, 108, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f),
        ("ref_moved_192x108", 192, 108, 0.3f, 0.5f, -1.0f, 0.2f, 0.15f),
        ("ref_above_160x90", 160, 90, -2.0f, 2.5f, 3.0f, -0.4f, 0.5f),
        ("ref_down_160x90", 160, 90, 0.0f, 3.0f, 2.0f, 0.0f, 1.3f),
        ("ref_behind_160x90", 160, 90, 0.0f, 0.0f, 6.0f, 3.1f, 0.0f),
        ("ref_sky_160x90", 160, 90, 0.0f, 0.5f, 0.0f, 0.0f, -0.6f),
        ("ref_default_1280x720", 1280, 720, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f),
    };

    public static void Main(string[] args)
    {
        string outDir = args.Length > 0 ? args[0] : "golden_ref";
        Directory.CreateDirectory(outDir);
        const BindingFlags priv = BindingFlags.Instance | BindingFlags.NonPublic;
        FieldInfo fPos = typeof(RayTracer).GetField("_cameraPosition", priv) ?? throw new Exception("RayTracer._cameraPosition not found");
        FieldInfo fYaw = typeof(RayTracer).GetField("_yaw", priv) ?? throw new Exception("RayTracer._yaw not found");
        FieldInfo fPitch = typeof(RayTracer).GetField("_pitch", priv) ?? throw new Exception("RayTracer._pitch not found");
        var manifest = new StringBuilder();
        manifest.Append("{\n  \"generator\": \"tests/golden/ref/GenGolden.cs\",\n");
        manifest.Append($"  \"runtime\": \"{RuntimeInformation.FrameworkDescription}\",\n");
        manifest.Append($"  \"arch\": \"{RuntimeInformation.ProcessArchitecture}\",\n");
        manifest.Append($"  \"opentk\": \"{typeof(Vector3).Assembly.GetName().Version}\",\n  \"cases\": [\n");
        for (int i = 0; i < Cases.Length; i++)
        {
            var c = Cases[i];
            var screen = new Surface(c.w, c.h);
            var rt = new RayTracer(screen);
            fPos.SetValue(rt, new Vector3(c.px, c.py, c.pz));
            fYaw.SetValue(rt, c.yaw);
            fPitch.SetValue(rt, c.pitch);
            rt.Tick();                                                   // the reference's own frame loop, unmodified
            var bytes = new byte[screen.pixels.Length * 4];
            Buffer.BlockCopy(screen.pixels, 0, bytes, 0, bytes.Length);  // little-endian on every .NET 6 target
            File.WriteAllBytes(Path.Combine(outDir, c.name + ".bin"), bytes);
            manifest.Append($"    {{\"name\": \"{c.name}\", \"w\": {c.w}, \"h\": {c.h}, \"pos\": [{F(c.px)}, {F(c.py)}, {F(c.pz)}], " +
                            $"\"yaw\": {F(c.yaw)}, \"pitch\": {F(c.pitch)}}}{(i + 1 < Cases.Length ? "," : "")}\n");
            Console.WriteLine($"{c.name}: {c.w}x{c.h} written");
        }
        manifest.Append("  ]\n}\n");
        File.WriteAllText(Path.Combine(outDir, "manifest.json"), manifest.ToString());
        MathProbe(Path.Combine(outDir, "ref_math.bin"));
    }

    private static string F(float v) => v.ToString("R", System.Globalization.CultureInfo.InvariantCulture);

    // xorshift32 -> floats in [-range, range); mirrored bit for bit by tests/test_reference_fixtures.py::_probe_inputs
    private static uint _s = 0x9E3779B9u;
    private static float Next(float range)
    {
        _s ^= _s << 13; _s ^= _s >> 17; _s ^= _s << 5;
        return ((_s >> 8) * (1.0f / 16777216.0f) * 2.0f - 1.0f) * range;
    }

    // N records of 16 float32 each:
    //   [0..2] Vector3.Normalize(a)   [3] Vector3.Dot(a, b)   [4..6] Vector3.Cross(a, b)   [7] a.Normalized().Length
    //   [8] (float)Math.Pow((double)|a.x/range|, 8.0)   [9] (float)Math.Pow((double)|a.y/range|, 0.5)   [10] (float)Math.Sqrt((double)|a.z|)
    //   [11] Math.Max(a.x, 0f)   [12] (float)(int)(a.y * 1e3f)   [13] (float)(1.0 / Math.Pow((double)a.z, 2.0))
    //   [14] (a * b).X (component-wise product)   [15] (1f / a.x) * a.x
    // preceded by N probe-input records of 6 float32 (a, b) so that the consumer need not trust its own PRNG mirror.
    private static void MathProbe(string path)
    {
        const int n = 4096;
        using var w = new BinaryWriter(File.Create(path));
        w.Write(n);
        var a = new Vector3[n]; var b = new Vector3[n];
        for (int i = 0; i < n; i++)
        {
            float range = i % 3 == 0 ? 1.0f : (i % 3 == 1 ? 40.0f : 1000.0f);
            a[i] = new Vector3(Next(range), Next(range), Next(range));
            b[i] = new Vector3(Next(range), Next(range), Next(range));
            w.Write(a[i].X); w.Write(a[i].Y); w.Write(a[i].Z); w.Write(b[i].X); w.Write(b[i].Y); w.Write(b[i].Z);
        }
        for (int i = 0; i < n; i++)
        {
            float range = i % 3 == 0 ? 1.0f : (i % 3 == 1 ? 40.0f : 1000.0f);
            Vector3 nrm = Vector3.Normalize(a[i]);
            Vector3 cr = Vector3.Cross(a[i], b[i]);
            w.Write(nrm.X); w.Write(nrm.Y); w.Write(nrm.Z);
            w.Write(Vector3.Dot(a[i], b[i]));
            w.Write(cr.X); w.Write(cr.Y); w.Write(cr.Z);
            w.Write(a[i].Normalized().Length);
            w.Write((float)Math.Pow((double)Math.Abs(a[i].X / range), 8.0));
            w.Write((float)Math.Pow((double)Math.Abs(a[i].Y / range), 0.5));
            w.Write((float)Math.Sqrt((double)Math.Abs(a[i].Z)));
            w.Write(Math.Max(a[i].X, 0f));
            w.Write((float)(int)(a[i].Y * 1e3f));
            w.Write((float)(1.0 / Math.Pow((double)a[i].Z, 2.0)));
            w.Write((a[i] * b[i]).X);
            w.Write((1f / a[i].X) * a[i].X);
        }
        Console.WriteLine($"ref_math.bin: {n} probe records written");
    }
}
