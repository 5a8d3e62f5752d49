// wc_headless -- what WaterCubeApp::update does every frame (src/WaterCubeApp.cpp:76-103),
// without the window: build the default scene (or a scaled one), step it with a fixed frame
// time (quirk Q15: the app uses the wall clock), print conserved / statistical quantities.
//
//   wc_headless [--particles N] [--size S] [--grid G] [--steps K] [--device D] [--dump file]
//               [--gpus N]         z-slab decomposition over N devices of this box, driven by this
//                                  one process and thread (core::Fluid::devices); same results
//               [--stiffness K] [--viscosity MU] [--rest-density R]   per-step parameters
//               [--wall-particles [--wall-density RHO]] [--surface-tension SIGMA]
//                                  the report's future-work physics (wc_physics); off = reference
//               [--initial-only]   (write the initial lattice to --dump; no device needed)
//               [--checkpoint file]  write a resumable checkpoint (header + AoS) after the run
//               [--restore file]     start from a checkpoint instead of the initial lattice
//
// The statistics line is computed on the host from the downloaded buffer; its "device" object
// is wc_diagnose's on-device reduction of the same buffer (they must agree).
//
// Exit code 0 on success, 2 when the native layer reports an error (e.g. no sm_100 device:
// there is no CPU fallback).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "core/Fluid.h"
#include "core/Scene.h"

using namespace core;

int main(int argc, char** argv) {
    int n = 80000, grid = 21, steps = 100, device = 0, gpus = 1;
    float size = 1.0f, stiffness = -1.0f, viscosity = -1.0f, rest_density = -1.0f;
    float surface_tension = 0.0f, wall_density = 0.0f;
    bool wall_particles = false;
    const char* dump = nullptr;
    const char* checkpoint = nullptr;
    const char* restore = nullptr;
    bool initial_only = false;
    for (int i = 1; i < argc; i++) {
        auto next = [&](const char* flag) -> const char* {
            if (std::strcmp(argv[i], flag) != 0) return nullptr;
            if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", flag); std::exit(1); }
            return argv[++i];
        };
        if (const char* v = next("--particles")) n = std::atoi(v);
        else if (const char* v = next("--size")) size = (float)std::atof(v);
        else if (const char* v = next("--grid")) grid = std::atoi(v);
        else if (const char* v = next("--steps")) steps = std::atoi(v);
        else if (const char* v = next("--device")) device = std::atoi(v);
        else if (const char* v = next("--gpus")) gpus = std::atoi(v);
        else if (const char* v = next("--stiffness")) stiffness = (float)std::atof(v);
        else if (const char* v = next("--viscosity")) viscosity = (float)std::atof(v);
        else if (const char* v = next("--rest-density")) rest_density = (float)std::atof(v);
        else if (const char* v = next("--surface-tension")) surface_tension = (float)std::atof(v);
        else if (const char* v = next("--wall-density")) wall_density = (float)std::atof(v);
        else if (std::strcmp(argv[i], "--wall-particles") == 0) wall_particles = true;
        else if (const char* v = next("--dump")) dump = v;
        else if (const char* v = next("--checkpoint")) checkpoint = v;
        else if (const char* v = next("--restore")) restore = v;
        else if (std::strcmp(argv[i], "--initial-only") == 0) initial_only = true;
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 1; }
    }
    try {
        FluidRef fluid = Fluid::create("fluid")->numParticles(n)->size(size)->gridRes(grid)->device(device)->devices(gpus);
        if (stiffness >= 0) fluid->stiffness(stiffness);
        if (viscosity >= 0) fluid->viscosityCoefficient(viscosity);
        if (rest_density >= 0) fluid->restDensity(rest_density);
        if (wall_particles) fluid->wallParticles(true, wall_density);
        if (surface_tension > 0) fluid->surfaceTension(surface_tension);
        if (initial_only) {
            const std::vector<Particle>& init = fluid->initialParticles();
            FILE* f = dump ? std::fopen(dump, "wb") : nullptr;
            if (!f || std::fwrite(init.data(), sizeof(Particle), init.size(), f) != init.size()) {
                std::fprintf(stderr, "cannot write the initial particles (--dump file)\n");
                return 1;
            }
            std::fclose(f);
            return 0;
        }
        if (restore) {
            fluid->restoreCheckpoint(restore);  // configures from the file, then setup()
            n = fluid->numParticles();
        } else {
            fluid->setup();  // WaterCubeApp.cpp:59-62
        }
        SceneRef scene = Scene::create();  // WaterCubeApp.cpp:64-66
        scene->addObject(fluid);
        const double frame = 1.0 / 60.0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int s = 0; s < steps; s++) scene->update(frame);  // Scene::update -> Fluid::update
        const std::vector<Particle> ps = util::getParticles(fluid->particleBuffer1(), n);  // syncs
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double ke = 0, mom[3] = {0, 0, 0}, com[3] = {0, 0, 0}, vmax = 0;
        long bad = 0;
        const double m = fluid->particleMass();
        for (const Particle& p : ps) {
            const double v2 = (double)p.velocity.x * p.velocity.x + (double)p.velocity.y * p.velocity.y +
                              (double)p.velocity.z * p.velocity.z;
            ke += 0.5 * m * v2;
            mom[0] += m * p.velocity.x, mom[1] += m * p.velocity.y, mom[2] += m * p.velocity.z;
            com[0] += p.position.x, com[1] += p.position.y, com[2] += p.position.z;
            vmax = std::fmax(vmax, std::sqrt(v2));
            // particle.vert:36-46: what render modes 1/2 flag red
            const bool out = !(p.position.x >= 0 && p.position.x <= size && p.position.y >= 0 &&
                               p.position.y <= size && p.position.z >= 0 && p.position.z <= size);
            if (out || !std::isfinite(v2) || !(p.density > 0) || !std::isfinite(p.density)) bad++;
        }
        const wc_diagnostics dg = fluid->diagnostics(1);
        std::printf("{\"particles\": %d, \"steps\": %d, \"total_steps\": %llu, \"seconds\": %.6f, "
                    "\"updates_per_sec\": %.6g, "
                    "\"kinetic_energy\": %.9g, \"momentum\": [%.9g, %.9g, %.9g], "
                    "\"centre_of_mass\": [%.9g, %.9g, %.9g], \"max_speed\": %.9g, \"invalid\": %ld, "
                    "\"device\": {\"kinetic_energy\": %.9g, \"invalid\": %lld, \"at_speed_clamp\": %lld, "
                    "\"density_mean\": %.9g, \"max_cell_count\": %lld}}\n",
                    n, steps, (unsigned long long)fluid->stepCount(), secs, (double)n * steps / secs, ke,
                    mom[0], mom[1], mom[2], com[0] / n, com[1] / n, com[2] / n, vmax, bad,
                    dg.kinetic_energy, (long long)dg.invalid, (long long)dg.at_speed_clamp,
                    dg.density_mean, (long long)dg.max_cell_count);
        if (checkpoint) fluid->saveCheckpoint(checkpoint);
        if (dump) {  // checkpoint: the 32-byte AoS array of util::getParticles (util.cpp:42-63)
            FILE* f = std::fopen(dump, "wb");
            if (!f || std::fwrite(ps.data(), sizeof(Particle), ps.size(), f) != ps.size()) {
                std::fprintf(stderr, "cannot write %s\n", dump);
                return 1;
            }
            std::fclose(f);
        }
        return bad ? 3 : 0;
    } catch (const core::Error& e) {
        std::fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 2;
    }
}
