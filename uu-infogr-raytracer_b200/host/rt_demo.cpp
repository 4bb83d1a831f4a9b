// rt_demo — headless stand-in for the reference's window loop (template.cs:175-213): N ticks of the C++ host mirror,
// a few simulated key presses / mouse moves, last frame written as a binary PPM.
//   rt_demo out.ppm [width height [ticks [n_devices [raylog.csv]]]]
// raylog.csv: the rays of 500 pseudo-random pixels of the last frame (the sample the DEBUG_ENABLE overlay draws, RayTracer.cs:914-933).
#include <cstdio>
#include <cstdlib>

#include "raytracer_host.hpp"

int main(int argc, char** argv) {
    const char* out = argc > 1 ? argv[1] : "frame.ppm";
    int w = argc > 3 ? atoi(argv[2]) : 1280, h = argc > 3 ? atoi(argv[3]) : 720;    // template.cs:65
    int ticks = argc > 4 ? atoi(argv[4]) : 1, ndev = argc > 5 ? atoi(argv[5]) : 1;
    try {
        rthost::Surface screen(w, h);
        rthost::RayTracer app(screen, ndev);
        for (int t = 0; t < ticks; t++) {
            if (t > 0) { app.OnKeyPress(rthost::Keys::W); app.OnKeyPress(rthost::Keys::D); app.OnMouseMove(3.6f, 1.8f); }
            app.Tick();
            fprintf(stderr, "tick %d: kernel %.3f ms, d2h %.3f ms\n", t, app.last_stats.kernel_ms, app.last_stats.d2h_ms);
        }
        FILE* f = fopen(out, "wb");
        if (!f) { perror(out); return 2; }
        fprintf(f, "P6\n%d %d\n255\n", w, h);
        for (int32_t p : screen.pixels) {
            unsigned char rgb[3] = {(unsigned char)((p >> 16) & 255), (unsigned char)((p >> 8) & 255), (unsigned char)(p & 255)};
            fwrite(rgb, 1, 3, f);
        }
        fclose(f);
        if (argc > 6) {
            std::vector<uint32_t> px(500);                                              // debugNumRays :916
            uint32_t st = 12345u;
            for (auto& p : px) { st = st * 747796405u + 2891336453u; p = (st >> 8) % (uint32_t)(w * h); }
            std::vector<rt_ray_record> recs = app.TracedRays(px);
            FILE* g = fopen(argv[6], "w");
            if (!g) { perror(argv[6]); return 2; }
            fprintf(g, "pixel,kind,level,light,hit,distance,ox,oy,oz,hx,hy,hz\n");
            for (const rt_ray_record& r : recs)
                fprintf(g, "%u,%u,%u,%u,%d,%.9g,%.9g,%.9g,%.9g,%.9g,%.9g,%.9g\n", r.pixel, r.kind, r.level, r.light, r.hit, r.distance,
                        r.origin[0], r.origin[1], r.origin[2], r.hit_point[0], r.hit_point[1], r.hit_point[2]);
            fclose(g);
            fprintf(stderr, "ray log: %zu records of %zu pixels -> %s\n", recs.size(), px.size(), argv[6]);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "rt_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
