// core/Fluid.cpp -- see Fluid.h.  Reference: src/core/Fluid.cpp.
#include "./Fluid.h"

#include <cmath>

namespace core {

namespace {

// The portable counter-based generator of watercube_b200/scenes.py (the reference seeds
// libc rand(), which differs between MSVC and glibc: quirk Q16).  Same integer mix, same
// float conversion, so a scene generated here equals scenes.dam_break bit for bit.
inline float hash_u01(uint64_t counter, uint32_t seed) {
    uint32_t x = (uint32_t)((counter + (uint64_t)(uint32_t)(seed * 0x9E3779B9u)) & 0xFFFFFFFFull);
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return (float)(x >> 8) * 5.9604644775390625e-08f;  // 2^-24
}

inline int lattice_side(int n) {  // int(ceil(cbrt(n))), Fluid.cpp:111, in integers
    int d = (int)std::lround(std::cbrt((double)n));
    while ((long long)d * d * d < n) d++;
    while (d > 0 && (long long)(d - 1) * (d - 1) * (d - 1) >= n) d--;
    return d;
}

}  // namespace

Fluid::Fluid(const std::string& name)
    : BaseObject(name),
      num_particles_(80000),          // Fluid.cpp:12
      grid_res_(21),                  // :14
      render_mode_(0),
      device_(0),
      seed_(0),
      size_(1.0f),                    // :10
      particle_radius_(0.01f),        // :17
      viscosity_coefficient_(200.0f), // :19
      stiffness_(100.0f),             // :20
      rest_density_(500.0f),          // :18
      rest_pressure_(0.0f),           // :21
      gravity_strength_(900.0f),      // :15
      time_scale_(0.012f),            // :24
      position_(0.0f),
      camera_position_(0.0f),
      light_position_(0.0f),
      gravity_direction_(0.0f, -1.0f, 0.0f),  // :16
      has_mouse_ray_(false),
      user_particles_(false),
      steps_(0),
      time_(0.0),
      handle_(nullptr) {
    derived_ = wc_derived();
}

Fluid::~Fluid() {
    if (handle_) wc_destroy(handle_);
}

FluidRef Fluid::numParticles(int n) { num_particles_ = n; return shared_from_this(); }
FluidRef Fluid::gridRes(int r) { grid_res_ = r; return shared_from_this(); }
FluidRef Fluid::size(float s) { size_ = s; return shared_from_this(); }
FluidRef Fluid::particleRadius(float r) { particle_radius_ = r; return shared_from_this(); }
FluidRef Fluid::position(vec3 p) { position_ = p; return shared_from_this(); }
FluidRef Fluid::renderMode(int m) { render_mode_ = m; return shared_from_this(); }
FluidRef Fluid::device(int ordinal) { device_ = ordinal; return shared_from_this(); }
FluidRef Fluid::seed(uint32_t s) { seed_ = s; return shared_from_this(); }
FluidRef Fluid::viscosityCoefficient(float c) { viscosity_coefficient_ = c; return shared_from_this(); }
FluidRef Fluid::stiffness(float s) { stiffness_ = s; return shared_from_this(); }
FluidRef Fluid::restDensity(float d) { rest_density_ = d; return shared_from_this(); }
FluidRef Fluid::restPressure(float p) { rest_pressure_ = p; return shared_from_this(); }
FluidRef Fluid::gravityStrength(float g) { gravity_strength_ = g; return shared_from_this(); }
FluidRef Fluid::gravityDirection(vec3 d) { gravity_direction_ = d; return shared_from_this(); }

FluidRef Fluid::initialParticles(const std::vector<Particle>& particles) {
    initial_particles_ = particles;
    num_particles_ = (int)particles.size();
    user_particles_ = true;
    if (handle_) util::setParticles(particleBuffer1(), initial_particles_);
    return shared_from_this();
}

// Fluid.cpp:104-134: d = ceil(cbrt(N)); jittered lattice with spacing 1.75 * r in the origin
// corner, x fastest, z slowest; velocity, density and pressure zero (util.h:30).
void Fluid::generateInitialParticles() {
    const int n = num_particles_;
    const int d = lattice_side(n);
    const float distance = particle_radius_ * 1.75f;  // :110
    const float jitter = distance * 0.5f;             // :113
    const float half = jitter / 2.0f;                 // :114
    initial_particles_.assign((size_t)n, Particle());
    for (int idx = 0; idx < n; idx++) {
        const int ix = idx % d, iy = (idx / d) % d, iz = idx / (d * d);
        const float jx = hash_u01(3ull * (uint64_t)idx + 0, seed_) * jitter - half;
        const float jy = hash_u01(3ull * (uint64_t)idx + 1, seed_) * jitter - half;
        const float jz = hash_u01(3ull * (uint64_t)idx + 2, seed_) * jitter - half;
        Particle& p = initial_particles_[(size_t)idx];
        p.position = vec3((float)ix * distance + jx, (float)iy * distance + jy,
                          (float)iz * distance + jz);  // :129-130
    }
}

const std::vector<Particle>& Fluid::initialParticles() {
    if (!user_particles_ && (int)initial_particles_.size() != num_particles_)
        generateInitialParticles();
    return initial_particles_;
}

FluidRef Fluid::setup() {
    util::log("creating fluid\n");
    if (handle_) {  // reset path (WaterCubeApp.cpp:81-88); the reference leaks its old buffers
        wc_destroy(handle_);
        handle_ = nullptr;
    }
    if (!user_particles_) generateInitialParticles();

    wc_params p;
    util::check(wc_default_params(&p));
    p.num_particles = num_particles_;
    p.grid_res = grid_res_;
    p.size = size_;
    p.particle_radius = particle_radius_;
    p.time_scale = time_scale_;
    p.device = device_;
    util::check(wc_create(&p, &handle_));         // prepareBuffers + compileShaders
    util::check(wc_get_derived(handle_, &derived_));
    util::log("bins %d, bin size %f, kernel radius %f, particle mass %f\n", derived_.num_bins,
              derived_.bin_size, derived_.kernel_radius, derived_.particle_mass);  // :211

    util::setParticles(particleBuffer1(), initial_particles_);  // :142-149

    // :229-231 -- the sorter works on this solver's buffers
    sort_ = Sort::create()->numItems(num_particles_)->gridRes(grid_res_)->binSize(derived_.bin_size);
    sort_->attach(handle_);
    sort_->prepareBuffers();
    sort_->compileShaders();
    steps_ = 0;
    time_ = 0.0;
    util::log("fluid created\n");
    return shared_from_this();
}

void Fluid::saveCheckpoint(const std::string& path) {
    if (!handle_) throw Error(WC_ERR_INVALID, "Fluid::saveCheckpoint before setup()");
    util::CheckpointHeader h = util::CheckpointHeader();
    h.grid_res = grid_res_;
    h.size = size_;
    h.particle_radius = particle_radius_;
    h.time_scale = time_scale_;
    h.steps = steps_;
    h.time = time_;
    util::saveCheckpoint(path, h, util::getParticles(particleBuffer1(), num_particles_));
}

FluidRef Fluid::restoreCheckpoint(const std::string& path) {
    util::CheckpointHeader h;
    std::vector<Particle> particles = util::loadCheckpoint(path, &h);
    grid_res_ = h.grid_res;
    size_ = h.size;
    particle_radius_ = h.particle_radius;
    time_scale_ = h.time_scale;
    initial_particles_.swap(particles);
    num_particles_ = (int)initial_particles_.size();
    user_particles_ = true;
    setup();
    steps_ = h.steps;
    time_ = h.time;
    return shared_from_this();
}

wc_diagnostics Fluid::diagnostics(int which) {
    if (!handle_) throw Error(WC_ERR_INVALID, "Fluid::diagnostics before setup()");
    wc_diagnostics d;
    util::check(wc_diagnose(handle_, which, rest_density_, &d));
    return d;
}

void Fluid::reset() { setup(); }

Ray Fluid::getRelativeMouseRay() const {
    // Fluid.cpp:251-256: the ray is translated by -position_.  Before the first mouse move
    // the reference's ray is uninitialised (quirk Q19); use one that misses the box.
    if (!has_mouse_ray_) return Ray(vec3(-10.0f, -10.0f, -10.0f), vec3(-1.0f, 0.0f, 0.0f));
    return Ray(mouse_ray_.getOrigin() - position_, mouse_ray_.getDirection());
}

wc_step_params Fluid::stepParams() const {
    wc_step_params sp;
    util::check(wc_default_step_params(&sp));
    sp.viscosity_coefficient = viscosity_coefficient_;
    sp.stiffness = stiffness_;
    sp.rest_density = rest_density_;
    sp.rest_pressure = rest_pressure_;
    const vec3 g = gravity_direction_ * gravity_strength_;  // Fluid.cpp:310
    sp.gravity[0] = g.x, sp.gravity[1] = g.y, sp.gravity[2] = g.z;
    const Ray ray = getRelativeMouseRay();
    sp.mouse_origin[0] = ray.origin.x, sp.mouse_origin[1] = ray.origin.y, sp.mouse_origin[2] = ray.origin.z;
    sp.mouse_dir[0] = ray.direction.x, sp.mouse_dir[1] = ray.direction.y, sp.mouse_dir[2] = ray.direction.z;
    return sp;
}

void Fluid::runDensityProg(Buffer particle_buffer) {
    if (particle_buffer.kind != BufferKind::Particles2)
        throw Error(WC_ERR_INVALID, "runDensityProg runs on the sorted buffer (Fluid.cpp:349)");
    const wc_step_params sp = stepParams();
    util::check(wc_density_only(handle_, &sp));
}

void Fluid::runUpdateProg(Buffer in_particles, Buffer out_particles, float time_step) {
    if (in_particles.kind != BufferKind::Particles2 || out_particles.kind != BufferKind::Particles1)
        throw Error(WC_ERR_INVALID, "runUpdateProg reads buffer 2 and writes buffer 1 (Fluid.cpp:350)");
    const wc_step_params sp = stepParams();
    util::check(wc_update_only(handle_, time_step, &sp));
}

// Fluid.cpp:342-354.  One wc_step is sort(buf1 -> buf2); density(buf2); update(buf2 -> buf1)
// enqueued back to back on the solver's stream; it returns without waiting for the GPU,
// like the reference's dispatches.  Readbacks (util::getParticles) synchronise.
void Fluid::update(double time) {
    if (!handle_) throw Error(WC_ERR_INVALID, "Fluid::update before setup()");
    const wc_step_params sp = stepParams();
    util::check(wc_step(handle_, (float)time, &sp));
    steps_++;
    time_ += time;
}

}  // namespace core
