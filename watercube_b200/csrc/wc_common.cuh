// wc_common.cuh -- shared device helpers and the per-launch constant block.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wc {

constexpr int kWarp = 32;

// Everything the density / update kernels read as GLSL uniforms
// (src/core/Fluid.cpp:276-285, :305-317), plus values precomputed once on the host.
struct SphConsts {
    int n;           // numParticles (targets of this launch)
    int first;       // index of the first target in the candidate arrays (0 unless slab mode)
    int G;           // gridRes
    int Gz;          // z-layers covered by the offsets table (G, or slab layers + 2 ghost layers)
    int zbase;       // global z-layer of the table's layer 0 (0, or slab z_begin - 1)
    float bin;       // binSize
    float size;      // size
    float h;         // kernelRadius
    float h2;        // h*h (fp32 product, as poly6Kernel forms it)
    float T;         // smallest fp32 x with sqrtf(x) >= h: "dist >= h" <=> d2 >= T
    float m;         // particleMass
    float poly6C, spikyC, viscC;
    float mu, k, rho0, P0;
    float g[3];
    float dt;
    float mo[3], md[3];
    int mouse_hits;  // update.comp:118-121 evaluated on the host (all-uniform expression)
};

// Extended physics (include/wc_sph.h: wc_physics, WC_PHYS_*).  Its own parameter type, taken only
// by the kernel instantiations that know the flags: the reference step's kernels keep their
// parameter layout (and with it their code, to the instruction).
struct SphConstsExt : SphConsts {
    uint32_t phys;   // WC_PHYS_* (never 0 in a launch of an extended instantiation ... or all off)
    float sigma;     // surface tension coefficient
    float n_min;     // colour-field gradient length above which a particle is "surface"
    float grad_m;    // -6 * poly6C * m: common factor of the colour-field sums
    float h2x3;      // 3 * h * h
    float wall_acc;  // wall_stiffness / dt^2
    float wall_d;    // wall_distance
    float wall_w;    // wall_rest_density * pi/4 * poly6C * h^9 (scale of the wall weight function)
};
template <bool kExt>
struct ConstsOfT { typedef SphConsts type; };
template <>
struct ConstsOfT<true> { typedef SphConstsExt type; };
template <bool kExt>
using ConstsOf = typename ConstsOfT<kExt>::type;
constexpr uint32_t kPhysWall = 1u, kPhysTension = 2u;

// ---------------------------------------------------------------------------------------
// z-slab mode (wc_slab.cuh).  Everything a step learns about itself -- how many particles the
// slab owns after this step's sort, how many sit in its first / last layer, how many ghosts
// the neighbours sent -- lives in this device record, written by the step's own kernels and
// read by the later ones, so that no launch parameter depends on a host read-back and the
// host never has to wait inside a step.
struct SlabDyn {
    uint32_t n;          // owned particles of the current step
    uint32_t n_first;    // ... of them in the first owned layer (the halo sent down)
    uint32_t n_last;     // ... in the last owned layer (the halo sent up)
    uint32_t n_glow;     // ghost particles received from below / above
    uint32_t n_ghigh;
    uint32_t peer_n[2];  // the neighbours' owned counts (where this rank's halo lands there)
    uint32_t n_in_old;   // owned count of the step's INPUT (= n of the previous step)
    uint32_t m_in[2];    // migrants received from below / above
    uint32_t errors;     // sticky: kSlabErr* bits
    uint32_t pad;
};
enum : uint32_t {
    kSlabErrOwned = 1u,      // more owned particles than `capacity`
    kSlabErrGhost = 2u,      // a received halo layer exceeds slab_ghost_capacity
    kSlabErrHalo = 4u,       // this rank's boundary layer exceeds the neighbours' ghost capacity
    kSlabErrMigrants = 8u,   // more migrants through one face than slab_migrant_capacity
    kSlabErrStray = 16u,     // a received migrant does not belong here (moved > 1 layer)
    kSlabErrTimeout = 32u,   // a neighbour's signal did not arrive (dead or out-of-order peer)
};

// What a slab kernel needs beyond its own arrays; all null / zero for a whole-grid handle.
struct SlabRef {
    SlabDyn* dyn;
    float4* peer_pos[2];       // the neighbours' buffer 2 (mapped peer memory): [0] below, [1] above
    float4* peer_vel[2];
    const uint32_t* wait[2];   // local flags the kernel waits on before it reads halo data
    uint32_t* raise[2];        // the neighbours' flags the kernel raises once ALL its blocks are done
    uint32_t* done;            // block counter of that kernel (zero at launch)
    uint32_t step_no;          // the value waited for / raised
    uint32_t Cg;               // ghost slots before the owned region of buffer 2
    uint32_t cap;              // capacity of the owned region
};

// Where this rank's first / last owned layer lives in the neighbours' buffer 2 (their ghost
// slots).  The kernels that PRODUCE halo data store it there as they go -- the reorder writes
// positions and velocities, the density pass adds density and pressure -- so the halo crosses
// NVLink under the producing kernel and no copy sits between the phases.
struct PeerHalo {
    float4* pos[2];     // [0] = the rank below, [1] = the rank above
    float4* vel[2];
    uint32_t dst[2];    // index of this rank's first halo particle in that neighbour's arrays
    uint32_t n_first;   // owned sorted particles [0, n_first) are the lower halo
    uint32_t hi_begin;  // owned sorted particles [hi_begin, n) are the upper halo
};

__device__ __forceinline__ PeerHalo peer_halo_of(const SlabRef& s) {
    PeerHalo ph;
    ph.pos[0] = s.peer_pos[0], ph.pos[1] = s.peer_pos[1];
    ph.vel[0] = s.peer_vel[0], ph.vel[1] = s.peer_vel[1];
    ph.dst[0] = ph.dst[1] = ph.n_first = ph.hi_begin = 0u;
    if (s.dyn && (s.peer_pos[0] || s.peer_pos[1])) {
        const SlabDyn d = *s.dyn;
        ph.n_first = d.n_first;
        ph.hi_begin = d.n - d.n_last;
        // sent down: the lower neighbour's ghost-high slice, right after its owned particles;
        // sent up: the upper neighbour's ghost-low slice, right before them
        ph.dst[0] = s.Cg + d.peer_n[0];
        ph.dst[1] = s.Cg - d.n_last;
    }
    return ph;
}

// A step that overflowed a capacity is dead: its later kernels do no work (every index they
// would derive from the counts could leave the buffers) but still raise their signals, and the
// host finds the sticky error word with the next info read.
__device__ __forceinline__ bool slab_dead(const SlabRef& s) { return s.dyn && s.dyn->errors != 0u; }

// ---- put-with-signal over NVLink --------------------------------------------------------
// One thread spins on a local flag until the neighbour raised it to step_no; bounded, so a dead
// or out-of-order neighbour ends in kSlabErrTimeout instead of a hang.
__device__ __forceinline__ void slab_spin(const SlabRef& s, const uint32_t* f) {
    if (!f || (s.dyn->errors & kSlabErrTimeout)) return;
    uint32_t v = 0;
    for (unsigned spins = 0;; spins++) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v >= s.step_no) break;
        if (spins > (1u << 26)) {  // ~ 20 s of 300 ns naps
            atomicOr(&s.dyn->errors, kSlabErrTimeout);
            break;
        }
        __nanosleep(300);
    }
    __threadfence_system();
}

// Block-level wait for both neighbours.  Every thread of the block must call it.
// (No indexing of s.wait by a thread-dependent value: the kernel parameter would be copied
// into local memory for it.)
__device__ __forceinline__ void slab_block_wait(const SlabRef& s) {
    if (!s.wait[0] && !s.wait[1]) return;
    if (threadIdx.x < 2) slab_spin(s, threadIdx.x == 0 ? s.wait[0] : s.wait[1]);
    __syncthreads();
}

// Warp-level wait for the neighbour below and / or above: the gathers call it only for the
// groups of the first / last owned layer, the only ones that read ghost slots -- every other
// warp starts at once and works while the halo is still on its way.  Every lane must call it
// with the same arguments.  The arguments pass through a vote, all lanes poll together and the
// loop ends on a vote: ptxas can then see that the control flow is warp-uniform.  With one
// polling lane (or a condition it cannot prove uniform) it loses its proof that the warp is
// converged for the REST of the kernel and wraps every ballot / shuffle of the gather loop in a
// convergence check (+6 % instructions in k_density_tile, profiles/r02_variant_sweep.md).
__device__ __forceinline__ void slab_warp_wait(const SlabRef& s, bool below, bool above) {
    const uint32_t* f0 = __any_sync(0xffffffffu, below) ? s.wait[0] : nullptr;
    const uint32_t* f1 = __any_sync(0xffffffffu, above) ? s.wait[1] : nullptr;
    if (!f0 && !f1) return;
    // The polls are relaxed loads and ONE acquire fence follows: an acquire load is a load plus
    // an invalidation of the SM's whole L1 (CCTL.IVALL), and __threadfence_system() a
    // sequentially consistent fence plus another -- three L1 flushes per boundary warp took
    // the other warps' candidate lines with them (the density pass lives on 52 % L1 hits).
    for (unsigned spins = 0;; spins++) {
        uint32_t v0 = s.step_no, v1 = s.step_no;
        if (f0) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v0) : "l"(f0) : "memory");
        if (f1) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v1) : "l"(f1) : "memory");
        const bool timed_out = spins > (1u << 26) || (s.dyn->errors & kSlabErrTimeout);
        if (__all_sync(0xffffffffu, (v0 >= s.step_no && v1 >= s.step_no) || timed_out)) {
            if (timed_out && (threadIdx.x & 31) == 0) atomicOr(&s.dyn->errors, kSlabErrTimeout);
            break;
        }
        __nanosleep(300);
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");  // pairs with the neighbour's st.release.sys
}

// Grid-level signal: of the `blocks` blocks that take part, the one that finishes LAST raises
// the flags at both neighbours.  Every thread of every participating block must call it, after
// its last store into a neighbour's memory; wrote_remote says whether this thread made one
// (a block without any skips the system-scope fence).  (No spin loop in here: a kernel that
// holds one under a lane-dependent condition loses ptxas' proof that its warps are converged,
// and every ballot / shuffle of the gather loop is then wrapped in a convergence check --
// +6 % instructions in k_density_tile, profiles/r02_variant_sweep.md.)
__device__ __forceinline__ void slab_grid_signal(const SlabRef& s, bool wrote_remote, uint32_t blocks) {
    if (!s.done) return;
    const int any_remote = __syncthreads_or(wrote_remote ? 1 : 0);
    if (threadIdx.x == 0) {
        // This block's remote stores (ordered by the barrier) before its ticket: a RELEASE at
        // system scope on the ticket itself.  (__threadfence_system() here is a sequentially
        // consistent fence plus an invalidation of the SM's L1, paid by every block of the
        // boundary layers while the other blocks of the SM are in their gather loops.)
        uint32_t ticket;
        if (any_remote)
            asm volatile("atom.release.sys.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(s.done) : "memory");
        else
            ticket = atomicAdd(s.done, 1u);
        if (ticket == blocks - 1u) {
            // acquire what the other blocks released (the tickets form one RMW chain), then
            // publish to the neighbours
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            if (s.raise[0])
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(s.raise[0]), "r"(s.step_no) : "memory");
            if (s.raise[1])
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(s.raise[1]), "r"(s.step_no) : "memory");
        }
    }
}

// count.comp:32, one component: clamp(int(p / binSize), 0, gridRes - 1) with an IEEE
// fp32 divide and truncation toward zero.  Same float-side clamp as the oracle
// (oracle/wc_oracle.cpp cell_coord) so NaN / huge inputs are defined identically.
__device__ __forceinline__ int cell_coord(float p, float bin, int G) {
    const float q = __fdiv_rn(p, bin);
    if (!(q >= 1.0f)) return 0;
    if (q >= (float)G) return G - 1;
    return __float2int_rz(q);
}

// count.comp:33.  zbase shifts the z-layer for a slab-local table (0 for the whole grid).
__device__ __forceinline__ uint32_t cell_index(float x, float y, float z, float bin, int G,
                                               int zbase = 0) {
    const uint32_t cx = (uint32_t)cell_coord(x, bin, G);
    const uint32_t cy = (uint32_t)cell_coord(y, bin, G);
    const uint32_t cz = (uint32_t)(cell_coord(z, bin, G) - zbase);
    return (cz * (uint32_t)G + cy) * (uint32_t)G + cx;
}

// Squared distance with the op order the oracle pins: fma(rz,rz, fma(ry,ry, rx*rx)).
__device__ __forceinline__ float dist2(float rx, float ry, float rz) {
    return __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, __fmul_rn(rx, rx)));
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace wc
