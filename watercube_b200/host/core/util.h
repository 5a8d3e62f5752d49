// core/util.h -- shared host types of the C++ facade.
//
// Mirrors the names of the reference's src/core/util.h (struct Particle :29-35,
// WORK_GROUP_SIZE :18, util::getParticles / setParticles / getUints / printParticles
// :47-67) without Cinder, OpenGL or <Windows.h>.  A GLuint buffer name becomes a
// `Buffer`: (handle of the native solver, which array).  Everything below talks to the
// device only through the C-ABI of include/wc_sph.h.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/wc_sph.h"

namespace core {

const int WORK_GROUP_SIZE = 128;  // util.h:18; kept for callers that size their own work

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }

// Stand-in for ci::Ray (origin + direction), Fluid::setMouseRay (Fluid.h:52).
struct Ray {
    vec3 origin, direction;
    Ray() {}
    Ray(vec3 o, vec3 d) : origin(o), direction(d) {}
    vec3 getOrigin() const { return origin; }
    vec3 getDirection() const { return direction; }
};

// Particle representation: identical 32-byte layout (util.h:29-35; std430 mirror in
// density.comp:5-10), so a reference-side std::vector<Particle> can be handed over as is.
struct Particle {
    Particle() : position(0), density(0), velocity(0), pressure(0) {}
    vec3 position;
    float density;
    vec3 velocity;
    float pressure;
};
static_assert(sizeof(Particle) == 32 && sizeof(Particle) == sizeof(wc_particle),
              "Particle must stay the 32-byte AoS record of src/core/util.h:29-35");

// What replaces a GLuint buffer name.
enum class BufferKind : int {
    None = 0,
    Particles1 = 1,  // particle_buffer1_: current state (Fluid.cpp:394 renders it)
    Particles2 = 2,  // particle_buffer2_: cell-sorted input with density / pressure
    Counts = 3,      // Sort::getCountBuffer
    Offsets = 4,     // Sort::getOffsetBuffer
    Sorted = 5,      // Sort::getSortedBuffer (sort.comp: sorted[dst] = particleID)
    CellIds = 6,     // per input particle: clamp(ivec3(p / binSize)) linearised (count.comp:32-33)
    NeighbourCounts = 7
};

struct Buffer {
    wc_handle* handle = nullptr;
    BufferKind kind = BufferKind::None;
    // A fluid decomposed over several devices (Fluid::devices): the z-slabs' handles in z
    // order; `handle` is then the first of them.  Reading such a buffer concatenates the
    // slabs, which IS the whole-grid array in the whole-grid order.
    const std::vector<wc_handle*>* slabs = nullptr;
    explicit operator bool() const { return handle != nullptr && kind != BufferKind::None; }
};

// Every failure of the native layer surfaces as this exception (the reference has no
// per-step error reporting at all; CI_ASSERT at setup only, Fluid.cpp:223).
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

namespace util {

void check(int rc);  // throws core::Error carrying wc_last_error()

void log(const char* format, ...);  // util.cpp:9-27, to stderr instead of OutputDebugStringA

std::vector<Particle> getParticles(Buffer buffer, int num_items);        // util.cpp:51-57
void setParticles(Buffer buffer, const std::vector<Particle>& particles);  // util.cpp:59-63
std::vector<uint32_t> getUints(Buffer buffer, int num_items);             // util.cpp:65-71

// util.cpp:113-126: position, velocity, density, pressure, ivec3(position / binSize).
void printParticles(Buffer particle_buffer, int n, float bin_size);

// Checkpoint file (version 2): a 160-byte header followed by the 32-byte AoS records exactly
// as getParticles returns them (util.cpp:42-63 layout), so `numpy.fromfile(f, float32,
// offset=160).reshape(-1, 8)` reads it.  Buffer 1 keeps its order across save / restore and the
// header carries every setup-time AND per-step parameter, so a restored run continues
// bit-identically without the caller re-applying anything (the stable sort only sees
// positions and order).  Version-1 files (64-byte header, no step parameters) still load.
struct CheckpointHeader {
    char magic[8];          // "WCB200\0\0"
    uint32_t version;       // 2
    int32_t num_particles;
    int32_t grid_res;
    float size;
    float particle_radius;
    float time_scale;
    uint64_t steps;         // Fluid::update calls that produced this state
    double time;            // sum of the frame times passed to update()
    uint8_t reserved[16];
    // ---- version 2: the wc_step_params inputs (Fluid.cpp:15-21) and the mouse ray
    float viscosity_coefficient, stiffness, rest_density, rest_pressure;
    float gravity_strength;
    float gravity_direction[3];
    float position[3];      // Fluid::position_, which the mouse ray is relative to
    int32_t has_mouse_ray;
    float mouse_origin[3], mouse_dir[3];
    // ---- the wc_physics record (extended physics).  All zero in files written before it
    // existed, which reads as "flags 0": the reference's step, the object's values kept.
    uint32_t physics_flags;
    float surface_tension, surface_threshold, wall_stiffness, wall_distance, wall_rest_density;
};
static_assert(sizeof(CheckpointHeader) == 160, "checkpoint header is 160 bytes");
constexpr size_t kCheckpointHeaderV1Bytes = 64;

void saveCheckpoint(const std::string& path, const CheckpointHeader& header,
                    const std::vector<Particle>& particles);
std::vector<Particle> loadCheckpoint(const std::string& path, CheckpointHeader* header);

}  // namespace util
}  // namespace core
