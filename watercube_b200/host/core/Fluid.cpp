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
      num_devices_(1),
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
    wc_default_physics(&physics_);
}

Fluid::~Fluid() { destroyHandles(); }

void Fluid::destroyHandles() {
    if (!slabs_.empty()) {
        // every slab first finishes what it has queued: its kernels store into the neighbours
        for (wc_handle* s : slabs_) wc_sync(s);
        for (wc_handle* s : slabs_) wc_destroy(s);
        slabs_.clear();
    } else if (handle_) {
        wc_destroy(handle_);
    }
    handle_ = nullptr;
}

FluidRef Fluid::numParticles(int n) { num_particles_ = n; return shared_from_this(); }
FluidRef Fluid::gridRes(int r) { grid_res_ = r; return shared_from_this(); }
FluidRef Fluid::size(float s) { size_ = s; return shared_from_this(); }
FluidRef Fluid::particleRadius(float r) { particle_radius_ = r; return shared_from_this(); }
FluidRef Fluid::position(vec3 p) { position_ = p; return shared_from_this(); }
FluidRef Fluid::renderMode(int m) { render_mode_ = m; return shared_from_this(); }
FluidRef Fluid::device(int ordinal) { device_ = ordinal; return shared_from_this(); }
FluidRef Fluid::devices(int n) { num_devices_ = n < 1 ? 1 : n; return shared_from_this(); }

FluidRef Fluid::wallParticles(bool on, float wall_rest_density) {
    physics_.flags = on ? (physics_.flags | WC_PHYS_WALL_PARTICLES)
                        : (physics_.flags & ~WC_PHYS_WALL_PARTICLES);
    physics_.wall_rest_density = wall_rest_density;
    applyPhysics();
    return shared_from_this();
}

FluidRef Fluid::surfaceTension(float sigma, float threshold) {
    if (sigma > 0.0f) {
        physics_.flags |= WC_PHYS_SURFACE_TENSION;
        physics_.surface_tension = sigma;
        physics_.surface_threshold = threshold;
    } else {
        physics_.flags &= ~WC_PHYS_SURFACE_TENSION;
    }
    applyPhysics();
    return shared_from_this();
}

FluidRef Fluid::physics(const wc_physics& ph) {
    physics_ = ph;
    applyPhysics();
    return shared_from_this();
}

void Fluid::applyPhysics() {
    if (!slabs_.empty()) {
        for (wc_handle* s : slabs_) util::check(wc_set_physics(s, &physics_));
    } else if (handle_) {
        util::check(wc_set_physics(handle_, &physics_));
    }
}
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
    if (handle_ && !slabs_.empty()) return setup();  // decomposed: cut the new set into slabs
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
    destroyHandles();  // reset path (WaterCubeApp.cpp:81-88); the reference leaks its old buffers
    if (!user_particles_) generateInitialParticles();
    if (num_devices_ > 1) {
        setupSlabs();
        applyPhysics();
        sort_.reset();
        steps_ = 0;
        time_ = 0.0;
        util::log("fluid created on %d slabs\n", (int)slabs_.size());
        return shared_from_this();
    }

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
    applyPhysics();
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

// The z-slab decomposition of SURVEY.md 8(e) behind the same setup(): equal-count cuts from the
// z-layer histogram of the initial particles (count.comp:32 on z), one slab handle per device,
// neighbours attached through peer memory (wc_slab_peer_attach).  Every slab keeps its
// particles in their input order, so per cell the slabs' inputs concatenate to the
// single-device input and the run is bit-identical to it.
void Fluid::setupSlabs() {
    wc_params p;
    util::check(wc_default_params(&p));
    p.num_particles = 0;
    p.grid_res = grid_res_;
    p.size = size_;
    p.particle_radius = particle_radius_;
    p.time_scale = time_scale_;
    util::check(wc_derive(&p, &derived_));
    int32_t visible = 0;
    util::check(wc_device_count(&visible));
    const int G = grid_res_, world = num_devices_;
    if (world > G) throw Error(WC_ERR_INVALID, "more slabs than z-layers");
    // global z-layer of every particle: clamp(int(z / binSize), 0, G - 1), IEEE divide (Q11/Q12)
    const float bin = derived_.bin_size;
    auto layer_of = [&](float z) {
        const float q = z / bin;
        if (!(q >= 1.0f)) return 0;
        if (q >= (float)G) return G - 1;
        return (int)q;
    };
    std::vector<long long> hist((size_t)G, 0);
    std::vector<int> layer(initial_particles_.size());
    for (size_t i = 0; i < initial_particles_.size(); i++) hist[(size_t)(layer[i] = layer_of(initial_particles_[i].position.z))]++;
    // cuts: every slab >= 1 layer, particle counts as equal as whole layers allow
    std::vector<long long> cum((size_t)G + 1, 0);
    for (int z = 0; z < G; z++) cum[(size_t)z + 1] = cum[(size_t)z] + hist[(size_t)z];
    const long long total = cum[(size_t)G];
    cuts_.assign(1, 0);
    for (int r = 1; r < world; r++) {
        const double target = (double)total * r / world;
        int z = 0;
        while (z < G && (double)cum[(size_t)z] < target) z++;  // first z with cum[z] >= target
        if (z > 0 && std::fabs((double)cum[(size_t)z - 1] - target) <= std::fabs((double)cum[(size_t)(z < G ? z : G)] - target)) z--;
        if (z < cuts_.back() + 1) z = cuts_.back() + 1;
        if (z > G - (world - r)) z = G - (world - r);
        cuts_.push_back(z);
    }
    cuts_.push_back(G);
    long long layer_max = 1;
    for (long long c : hist) layer_max = c > layer_max ? c : layer_max;
    std::vector<int> rank_of_layer((size_t)G, 0);
    for (int r = 0; r < world; r++)
        for (int z = cuts_[(size_t)r]; z < cuts_[(size_t)r + 1]; z++) rank_of_layer[(size_t)z] = r;
    std::vector<std::vector<Particle>> parts((size_t)world);
    for (size_t i = 0; i < initial_particles_.size(); i++)
        parts[(size_t)rank_of_layer[(size_t)layer[i]]].push_back(initial_particles_[i]);
    try {
        for (int r = 0; r < world; r++) {
            wc_params q = p;
            q.device = device_ + (r % visible);
            const long long mine = (long long)parts[(size_t)r].size();
            q.capacity = (int32_t)(mine + mine / 4 + 4 * layer_max + 1024);
            q.slab_z_begin = cuts_[(size_t)r];
            q.slab_z_end = cuts_[(size_t)r + 1];
            q.slab_ghost_capacity = (int32_t)(layer_max + layer_max / 2 + 1024);
            q.slab_migrant_capacity = (int32_t)(layer_max / 2 > 65536 ? layer_max / 2 : 65536);
            wc_handle* h = nullptr;
            util::check(wc_create(&q, &h));
            slabs_.push_back(h);
        }
        for (int r = 0; r < world; r++) {
            if (r > 0) util::check(wc_slab_peer_attach(slabs_[(size_t)r], 0, slabs_[(size_t)r - 1]));
            else util::check(wc_slab_clear_recv(slabs_[(size_t)r], 0));
            if (r + 1 < world) util::check(wc_slab_peer_attach(slabs_[(size_t)r], 1, slabs_[(size_t)r + 1]));
            else util::check(wc_slab_clear_recv(slabs_[(size_t)r], 1));
            util::check(wc_upload_particles(slabs_[(size_t)r],
                                            reinterpret_cast<const wc_particle*>(parts[(size_t)r].data()),
                                            (int32_t)parts[(size_t)r].size()));
        }
    } catch (...) {
        for (wc_handle* s : slabs_) wc_destroy(s);
        slabs_.clear();
        throw;
    }
    handle_ = slabs_[0];
    util::log("bins %d, bin size %f, kernel radius %f, particle mass %f; %d z-slabs\n",
              derived_.num_bins, derived_.bin_size, derived_.kernel_radius, derived_.particle_mass, world);
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
    h.viscosity_coefficient = viscosity_coefficient_;
    h.stiffness = stiffness_;
    h.rest_density = rest_density_;
    h.rest_pressure = rest_pressure_;
    h.gravity_strength = gravity_strength_;
    h.gravity_direction[0] = gravity_direction_.x, h.gravity_direction[1] = gravity_direction_.y,
    h.gravity_direction[2] = gravity_direction_.z;
    h.position[0] = position_.x, h.position[1] = position_.y, h.position[2] = position_.z;
    h.has_mouse_ray = has_mouse_ray_ ? 1 : 0;
    const vec3 mo = mouse_ray_.getOrigin(), md = mouse_ray_.getDirection();
    h.mouse_origin[0] = mo.x, h.mouse_origin[1] = mo.y, h.mouse_origin[2] = mo.z;
    h.mouse_dir[0] = md.x, h.mouse_dir[1] = md.y, h.mouse_dir[2] = md.z;
    h.physics_flags = physics_.flags;
    h.surface_tension = physics_.surface_tension, h.surface_threshold = physics_.surface_threshold;
    h.wall_stiffness = physics_.wall_stiffness, h.wall_distance = physics_.wall_distance;
    h.wall_rest_density = physics_.wall_rest_density;
    util::saveCheckpoint(path, h, util::getParticles(particleBuffer1(), num_particles_));
}

FluidRef Fluid::restoreCheckpoint(const std::string& path) {
    util::CheckpointHeader h;
    std::vector<Particle> particles = util::loadCheckpoint(path, &h);
    grid_res_ = h.grid_res;
    size_ = h.size;
    particle_radius_ = h.particle_radius;
    time_scale_ = h.time_scale;
    if (h.version >= 2) {  // version 1 carried no step parameters: the object's stay in force
        viscosity_coefficient_ = h.viscosity_coefficient;
        stiffness_ = h.stiffness;
        rest_density_ = h.rest_density;
        rest_pressure_ = h.rest_pressure;
        gravity_strength_ = h.gravity_strength;
        gravity_direction_ = vec3(h.gravity_direction[0], h.gravity_direction[1], h.gravity_direction[2]);
        position_ = vec3(h.position[0], h.position[1], h.position[2]);
        has_mouse_ray_ = h.has_mouse_ray != 0;
        mouse_ray_ = Ray(vec3(h.mouse_origin[0], h.mouse_origin[1], h.mouse_origin[2]),
                         vec3(h.mouse_dir[0], h.mouse_dir[1], h.mouse_dir[2]));
        if (h.physics_flags != 0u) {  // (all zero in files written before the record existed)
            physics_.flags = h.physics_flags;
            physics_.surface_tension = h.surface_tension, physics_.surface_threshold = h.surface_threshold;
            physics_.wall_stiffness = h.wall_stiffness, physics_.wall_distance = h.wall_distance;
            physics_.wall_rest_density = h.wall_rest_density;
        } else {
            physics_.flags = 0u;
        }
    }
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
    if (slabs_.empty()) {
        util::check(wc_diagnose(handle_, which, rest_density_, &d));
        return d;
    }
    // decomposed: the slabs' on-device reductions folded in z order (sums add; the centre of
    // mass and the means are re-weighted by the slabs' valid particles)
    std::vector<wc_diagnostics> parts(slabs_.size());
    for (size_t r = 0; r < slabs_.size(); r++) util::check(wc_diagnose(slabs_[r], which, rest_density_, &parts[r]));
    d = wc_diagnostics();
    d.max_cell_count = d.nonempty_cells = -1;
    double valid = 0, rho_sum = 0, pres_sum = 0;
    bool first = true;
    const double m = (double)derived_.particle_mass;
    for (const wc_diagnostics& q : parts) {
        const double nv = m > 0 ? q.mass / m : 0.0;
        d.particles += q.particles, d.invalid += q.invalid, d.out_of_box += q.out_of_box;
        d.at_speed_clamp += q.at_speed_clamp;
        d.mass += q.mass, d.kinetic_energy += q.kinetic_energy;
        for (int a = 0; a < 3; a++) d.momentum[a] += q.momentum[a], d.centre_of_mass[a] += q.centre_of_mass[a] * nv;
        d.max_speed = q.max_speed > d.max_speed ? q.max_speed : d.max_speed;
        if (nv > 0) {
            d.density_min = first ? q.density_min : (q.density_min < d.density_min ? q.density_min : d.density_min);
            d.density_max = first ? q.density_max : (q.density_max > d.density_max ? q.density_max : d.density_max);
            d.pressure_min = first ? q.pressure_min : (q.pressure_min < d.pressure_min ? q.pressure_min : d.pressure_min);
            d.pressure_max = first ? q.pressure_max : (q.pressure_max > d.pressure_max ? q.pressure_max : d.pressure_max);
            first = false;
        }
        rho_sum += q.density_mean * nv, pres_sum += q.pressure_mean * nv, valid += nv;
        for (int k = 0; k < WC_DIAG_HIST_BINS; k++) d.density_hist[k] += q.density_hist[k];
    }
    if (valid > 0) {
        for (int a = 0; a < 3; a++) d.centre_of_mass[a] /= valid;
        d.density_mean = rho_sum / valid, d.pressure_mean = pres_sum / valid;
    }
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
    if (slabs_.empty()) {
        util::check(wc_step(handle_, (float)time, &sp));
    } else {
        // every slab's whole step, queued back to back by this one thread: nothing in a step
        // waits for the host (wc_slab_step_peer with info == NULL), and a kernel that waits for
        // a neighbour only waits for work that is queued here without anybody waiting for it
        for (wc_handle* s : slabs_) util::check(wc_slab_step_peer(s, (float)time, &sp, nullptr));
    }
    steps_++;
    time_ += time;
}

}  // namespace core
