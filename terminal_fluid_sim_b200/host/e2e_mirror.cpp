// e2e_mirror.cpp -- the frame loop of the reference (src/app/app.rs:69-87: next_step, then render) driven through
// include/fluid_sim.hpp, the C++ mirror of `impl FluidSim`, exactly as a host application would use it: a brush
// edit (set_block / unset_block), next_step(dt), the renderer's visible window (render_view).  Prints one JSON line
// with the wall-clock time per frame and the state checksum, so bench.py and the GPU tests can show that (a) the
// wrapper's next_step() does no whole-grid device-to-host copy and (b) it computes the same bits as the ctypes mirror.
//   e2e_mirror W H K gravity scene frames [mode]      scene: none | disc | lattice | random
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fluid_sim.hpp"

int main(int argc, char **argv) {
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s W H K gravity scene frames [mode]\n", argv[0]);
        return 2;
    }
    const size_t W = std::strtoull(argv[1], nullptr, 10), H = std::strtoull(argv[2], nullptr, 10);
    const int K = std::atoi(argv[3]);
    const float gravity = (float)std::atof(argv[4]);
    const char *scene = argv[5];
    const int frames = std::atoi(argv[6]);
    const int mode = argc > 7 ? std::atoi(argv[7]) : 0;
    const float dt = 1.0f / 60.0f;
    try {
        tfs_options opt;
        tfs_default_options(&opt);
        opt.iterations = K;
        opt.mode = mode;
        tfs::SimConfig cfg;
        cfg.gravity = gravity;
        tfs::FluidSim sim(W, H, cfg, opt);
        if (!std::strcmp(scene, "disc")) tfs_scene_disc(sim.handle(), 256, 512, 128);
        if (!std::strcmp(scene, "lattice")) tfs_scene_lattice(sim.handle(), 1024, 96);
        if (!std::strcmp(scene, "random")) tfs_scene_random(sim.handle(), 0x5EED, (uint64_t)(((unsigned __int128)2 << 64) / 100));
        const size_t aw = W < 512 ? W : 512, ah = (H < 256 ? H : 256) / 2;
        std::vector<uint8_t> cells;
        auto frame = [&](int i) {
            const size_t x = 8 + (size_t)(i % 8) < W ? 8 + (size_t)(i % 8) : W - 1;
            for (size_t y = H / 2 - (H >= 8 ? 4 : 0); y < H / 2 + (H >= 8 ? 4 : 0); ++y) {  // the brush
                if (i & 1) sim.unset_block(x, y); else sim.set_block(x, y);
            }
            sim.next_step(dt);
            sim.render_view(aw, ah, cells);  // synchronises: the frame is complete when this returns
        };
        for (int i = 0; i < 3; ++i) frame(i);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < frames; ++i) frame(3 + i);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t cs[8];
        if (tfs_checksum(sim.handle(), cs) != TFS_OK) throw std::runtime_error(tfs_last_error());
        uint64_t launches = 0;
        tfs_launch_count(sim.handle(), &launches);
        std::printf("{\"ms_per_frame\": %.4f, \"gcell_steps_per_s\": %.5f, \"frames\": %d, \"d2h_bytes_per_frame\": %zu, "
                    "\"h2d_bytes_per_frame\": %zu, \"launches\": %" PRIu64 ", \"checksum\": [",
                    s / frames * 1e3, (double)W * (double)H * frames / s / 1e9, frames, aw * ah * 8 + 8, (size_t)(H >= 8 ? 8 : 0),
                    launches);
        for (int i = 0; i < 8; ++i) std::printf("%s\"%016" PRIx64 "\"", i ? ", " : "", cs[i]);
        std::printf("]}\n");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "e2e_mirror: %s\n", e.what());
        return 1;
    }
    return 0;
}
