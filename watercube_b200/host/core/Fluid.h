// core/Fluid.h -- host facade of the SPH solver.
//
// Keeps the setup / update / particle-buffer surface of the reference's core::Fluid
// (src/core/Fluid.h:32-59) so that Scene::update (Scene.cpp:52-56) -- and with it the Cinder
// app's update call (WaterCubeApp.cpp:98) -- can drive the B200 path unchanged: same fluent
// setters, same setup(), same virtual update(double time).  All device work goes through the
// C-ABI of include/wc_sph.h; there is no CPU fallback (setup() throws core::Error when no
// sm_100 device is usable).  Rendering and the AntTweakBar panel (Fluid.cpp:89-99,
// :359-431) are out of scope: draw() is a no-op and the GUI-bound fields are plain members
// read on every update, exactly as the reference re-uploads them as uniforms every frame
// (Fluid.cpp:276-285, :305-317).
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "./BaseObject.h"
#include "./Sort.h"
#include "./util.h"

namespace core {

typedef std::shared_ptr<class Fluid> FluidRef;

class Fluid : public BaseObject, public std::enable_shared_from_this<Fluid> {
public:
    explicit Fluid(const std::string& name);
    ~Fluid();
    Fluid(const Fluid&) = delete;
    Fluid& operator=(const Fluid&) = delete;

    int numParticles() { return num_particles_; }
    // Setup-time setters (call before setup(), Fluid.cpp:31-84).  They return the object
    // itself rather than the reference's copies (quirk Q17).
    FluidRef numParticles(int n);
    FluidRef gridRes(int r);
    FluidRef size(float s);
    FluidRef particleRadius(float r);
    FluidRef position(vec3 p);
    FluidRef renderMode(int m);
    FluidRef device(int ordinal);            // new: CUDA device ordinal (the first one used)
    // new: decompose the fluid into n z-slabs, one per device device()..device()+n-1 of this
    // box: equal-count cuts from the initial particles, the neighbours attached
    // through peer memory, every update() queued on all of them by this one host thread with
    // no wait inside a step.  The particle-buffer surface stays: buffer 1 read back is the
    // slabs in z order = the whole-grid array, bit-identical to the single-device run.  With
    // fewer devices than slabs the slabs share devices round-robin (a test configuration).
    FluidRef devices(int n);
    FluidRef seed(uint32_t s);               // new: seed of the portable jitter generator (Q16)
    // Per-step mutable (GUI-bound in the reference).
    FluidRef viscosityCoefficient(float c);
    FluidRef stiffness(float s);
    FluidRef restDensity(float d);
    FluidRef restPressure(float p);
    FluidRef gravityStrength(float g);
    FluidRef gravityDirection(vec3 d);       // what "Rotate Gravity" updates (Fluid.cpp:259-263)
    // new: the physics the reference's report lists as future work (report.pdf section 6; see
    // wc_physics in wc_sph.h), off by default = the reference's step.  Before setup() or between
    // two update() calls; carried by checkpoints.
    FluidRef wallParticles(bool on, float wall_rest_density = 0.0f);   // Harada et al.'s walls
    FluidRef surfaceTension(float sigma, float threshold = 7.0f);      // sigma <= 0: off
    FluidRef physics(const wc_physics& ph);
    const wc_physics& physics() const { return physics_; }

    void setCameraPosition(vec3 p) { camera_position_ = p; }
    void setLightPosition(vec3 p) { light_position_ = p; }
    void setMouseRay(Ray r) { mouse_ray_ = r; has_mouse_ray_ = true; }

    // Replace the generated lattice with caller-provided particles (before or after setup()).
    FluidRef initialParticles(const std::vector<Particle>& particles);

    FluidRef setup();                    // Fluid.cpp:203-235
    void update(double time) override;   // Fluid.cpp:342-354
    void draw() override {}              // graphics, out of scope
    void reset() override;               // WaterCubeApp.cpp:81-88: rebuild from the initial state

    // The particle-buffer surface: buffer 1 = current state (what the renderer binds,
    // Fluid.cpp:394), buffer 2 = cell-sorted input of the last step with density / pressure.
    Buffer particleBuffer1() const { return buffer(BufferKind::Particles1); }
    Buffer particleBuffer2() const { return buffer(BufferKind::Particles2); }
    SortRef sort() const { return sort_; }  // single device only (null for a decomposed fluid)
    wc_handle* nativeHandle() const { return handle_; }
    const std::vector<wc_handle*>& slabHandles() const { return slabs_; }  // empty: single device
    const std::vector<int>& slabCuts() const { return cuts_; }             // z-layer cuts [0..G]

    // Checkpoint / resume: the state of buffer 1 with the setup-time parameters.
    // restoreCheckpoint() configures the object from the file and calls setup(); the run then
    // continues bit-identically to one that was never interrupted.
    void saveCheckpoint(const std::string& path);
    FluidRef restoreCheckpoint(const std::string& path);
    uint64_t stepCount() const { return steps_; }
    double simulatedTime() const { return time_; }
    // On-device inspection of buffer 1 / 2 (what render modes 1-4 show).
    wc_diagnostics diagnostics(int which = 1);

    // Derived constants of Fluid::setup (Fluid.cpp:206-216).
    float binSize() const { return derived_.bin_size; }
    float kernelRadius() const { return derived_.kernel_radius; }
    float particleMass() const { return derived_.particle_mass; }
    int numBins() const { return derived_.num_bins; }
    // The initial lattice (generated on first use; does not touch the device).
    const std::vector<Particle>& initialParticles();

    static FluidRef create(const std::string& name) { return std::make_shared<Fluid>(name); }

protected:
    void generateInitialParticles();  // Fluid.cpp:104-134 with a portable generator (Q16)
    Ray getRelativeMouseRay() const;  // Fluid.cpp:251-256
    void runDensityProg(Buffer particle_buffer);                              // Fluid.cpp:268
    void runUpdateProg(Buffer in_particles, Buffer out_particles, float time_step);  // :294
    wc_step_params stepParams() const;
    Buffer buffer(BufferKind kind) const {
        Buffer b{handle_, kind};
        if (!slabs_.empty()) b.slabs = &slabs_;
        return b;
    }
    void setupSlabs();   // the decomposed counterpart of setup()'s buffer creation
    void destroyHandles();
    void applyPhysics();  // physics_ into every handle there is

    int num_particles_, grid_res_, render_mode_, device_, num_devices_;
    std::vector<wc_handle*> slabs_;  // z order; handle_ == slabs_[0] when decomposed
    std::vector<int> cuts_;
    uint32_t seed_;
    float size_, particle_radius_, viscosity_coefficient_, stiffness_, rest_density_,
        rest_pressure_, gravity_strength_, time_scale_;
    vec3 position_, camera_position_, light_position_, gravity_direction_;
    Ray mouse_ray_;
    bool has_mouse_ray_, user_particles_;
    uint64_t steps_;
    double time_;
    std::vector<Particle> initial_particles_;
    SortRef sort_;
    wc_handle* handle_;
    wc_derived derived_;
    wc_physics physics_;
};

}  // namespace core
