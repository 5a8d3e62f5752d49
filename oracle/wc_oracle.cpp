// wc_oracle.cpp -- CPU oracle for the WaterCube SPH step.  TEST INFRASTRUCTURE ONLY
// (see wc_oracle.h for the scope rules and the "parity unpinned" statement).
//
// Every function cites the reference file:line it restates.  The GLSL sources are
// under /root/reference/assets; the host constants under /root/reference/src/core.
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp; never -ffast-math).

#include "wc_oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// density.comp:40-50 == update.comp:49-59: the 27 neighbour offsets, dx outermost,
// dz innermost (quirk Q13: visiting order differs from the memory order x-fastest).
struct I3 {
    int x, y, z;
};
const I3 NEIGHBORHOOD[27] = {
    {-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {-1, 1, -1},
    {-1, 1, 0},   {-1, 1, 1},  {0, -1, -1}, {0, -1, 0},  {0, -1, 1}, {0, 0, -1}, {0, 0, 0},
    {0, 0, 1},    {0, 1, -1},  {0, 1, 0},   {0, 1, 1},   {1, -1, -1}, {1, -1, 0}, {1, -1, 1},
    {1, 0, -1},   {1, 0, 0},   {1, 0, 1},   {1, 1, -1},  {1, 1, 0},  {1, 1, 1}};

// count.comp:32: clamp(ivec3(p / binSize), 0, gridRes-1) for one component.
// ivec3() truncates toward zero (Q11).  The clamp is done in float first so NaN and
// out-of-int-range quotients are defined (NaN -> 0, +huge -> G-1); for every finite
// quotient this equals trunc-then-clamp.
inline int cell_coord(float p, float bin_size, int grid_res) {
    const float q = p / bin_size;  // IEEE fp32 divide (Q12)
    if (!(q >= 1.0f)) return 0;
    if (q >= (float)grid_res) return grid_res - 1;
    return (int)q;
}

inline uint32_t cell_index(const float* pos, float bin_size, int G) {
    const int cx = cell_coord(pos[0], bin_size, G);
    const int cy = cell_coord(pos[1], bin_size, G);
    const int cz = cell_coord(pos[2], bin_size, G);
    // count.comp:33
    return (uint32_t)cz * (uint32_t)G * (uint32_t)G + (uint32_t)cy * (uint32_t)G + (uint32_t)cx;
}

// length(r) of density.comp:116 / update.comp:172, with the dot product's op order
// pinned: fma(rz,rz, fma(ry,ry, rx*rx)), then correctly rounded sqrt.
inline float dist2_f32(float rx, float ry, float rz) {
    return fmaf(rz, rz, fmaf(ry, ry, rx * rx));
}

template <typename T>
struct Consts {
    T size, bin, h, m, poly6C, spikyC, viscC;
    T mu, k, rho0, P0;
    T g[3];
    T mo[3], md[3];
    T eps;  // the literal 1e-16 of update.comp:131,181,198 as a float
    int G;
    // extended physics (wc_oracle.h: WCO_PHYS_*); phys == 0 is the reference's step
    uint32_t phys;
    T sigma, n_min, wall_k, wall_d, wall_rho;
    T wallC;  // pi/4 * poly6C * h^9: the half-space integral of W_poly6 is wallC * 128/315
};

template <typename T>
Consts<T> make_consts(const wco_params* p) {
    wco_derived d;
    wco_derive(p, &d);
    Consts<T> c;
    c.size = (T)p->size;
    c.bin = (T)d.bin_size;
    c.h = (T)d.kernel_radius;
    c.m = (T)d.particle_mass;
    c.poly6C = (T)d.poly6_const;
    c.spikyC = (T)d.spiky_const;
    c.viscC = (T)d.visc_const;
    c.mu = (T)p->viscosity_coefficient;
    c.k = (T)p->stiffness;
    c.rho0 = (T)p->rest_density;
    c.P0 = (T)p->rest_pressure;
    for (int a = 0; a < 3; a++) {
        c.g[a] = (T)p->gravity[a];
        c.mo[a] = (T)p->mouse_origin[a];
        c.md[a] = (T)p->mouse_dir[a];
    }
    c.eps = (T)1e-16f;
    c.G = p->grid_res;
    c.phys = p->physics_flags;
    c.sigma = (T)p->surface_tension;
    c.n_min = (T)p->surface_threshold;
    c.wall_k = (T)p->wall_stiffness;
    c.wall_d = (T)p->wall_distance;
    c.wall_rho = p->wall_rest_density > 0.0f ? (T)p->wall_rest_density : (T)p->rest_density;
    {   // one fp32 constant for both precisions (the CUDA host derives the very same float)
        const double h = (double)d.kernel_radius;
        c.wallC = (T)(float)(0.78539816339744830962 * (double)d.poly6_const * std::pow(h, 9.0));
    }
    return c;
}

// density.comp:53-55.  pow(x,3) := (x*x)*x (Q3).
template <typename T>
inline T poly6(const Consts<T>& c, T r) {
    const T t = c.h * c.h - r * r;
    return ((t * t) * t) * c.poly6C;
}

// density.comp:57-79, including the z-high branch that tests p.y (Q2) and therefore
// can feed poly6 an r > h (negative contribution, Q3).
template <typename T>
inline T wall_density(const Consts<T>& c, const T* p) {
    T density = 0;
    if (p[0] < c.h) {
        density += c.m * poly6(c, p[0]);
    } else if (p[0] > c.size - c.h) {
        density += c.m * poly6(c, c.size - p[0]);
    }
    if (p[1] < c.h) {
        density += c.m * poly6(c, p[1]);
    } else if (p[1] > c.size - c.h) {
        density += c.m * poly6(c, c.size - p[1]);
    }
    if (p[2] < c.h) {
        density += c.m * poly6(c, p[2]);
    } else if (p[1] > c.size - c.h) {  // sic: density.comp:74 tests p.y
        density += c.m * poly6(c, c.size - p[2]);
    }
    return density * (T)4;
}

// update.comp:62-64.  pow(x,2) := x*x (Q6); evaluation order (s*(r/d))*C.
template <typename T>
inline void spiky(const Consts<T>& c, const T* r, T d, T* out) {
    const T s = (c.h - d) * (c.h - d);
    for (int a = 0; a < 3; a++) out[a] = (s * (r[a] / d)) * c.spikyC;
}

// update.comp:67-69
template <typename T>
inline T wvis(const Consts<T>& c, T r) {
    return (c.h - r) * c.viscC;
}

// update.comp:71-100.  r.length() is the GLSL component count == 3 (Q5).
template <typename T>
inline void wall_forces(const Consts<T>& c, const T* p, T* force) {
    force[0] = force[1] = force[2] = 0;
    const T three = (T)3;
    for (int a = 0; a < 3; a++) {
        T r[3];
        bool hit = false;
        if (p[a] < c.h) {
            for (int b = 0; b < 3; b++) r[b] = (b == a ? (T)0 : p[b]) - p[b];
            hit = true;
        } else if (p[a] > c.size - c.h) {
            for (int b = 0; b < 3; b++) r[b] = (b == a ? c.size : p[b]) - p[b];
            hit = true;
        }
        if (hit) {
            T w[3];
            spiky(c, r, three, w);
            for (int b = 0; b < 3; b++) force[b] += w[b];
        }
    }
    for (int b = 0; b < 3; b++) force[b] = force[b] * (T)0.01f;
}

// --- extended physics (not in the reference; wc_oracle.h WCO_PHYS_WALL_PARTICLES) ----------
// Wall weight function: wall_rho * integral of W_poly6 over the half space at distance >= s,
//   wall_rho * wallC * (128/315 - P(u)),  u = s / h,  P(u) = u - 4u^3/3 + 6u^5/5 - 4u^7/7 + u^9/9
// (Horner in u^2, this operation order), summed over the walls closer than h.
template <typename T>
inline T wall_weight(const Consts<T>& c, T s) {
    const T u = s / c.h, u2 = u * u;
    T P = (T)(1.0f / 9.0f);
    P = P * u2 + (T)(-4.0f / 7.0f);
    P = P * u2 + (T)(6.0f / 5.0f);
    P = P * u2 + (T)(-4.0f / 3.0f);
    P = P * u2 + (T)1;
    P = P * u;
    return (c.wall_rho * c.wallC) * ((T)(128.0f / 315.0f) - P);
}
template <typename T>
inline T wall_density_particles(const Consts<T>& c, const T* p) {
    T density = 0;
    for (int a = 0; a < 3; a++) {
        if (p[a] < c.h) {
            density += wall_weight(c, p[a]);
        } else if (p[a] > c.size - c.h) {
            density += wall_weight(c, c.size - p[a]);
        }
    }
    return density;
}
// Penetration push: acceleration wall_k * (wall_d - s) / dt^2 along the inward normal.
template <typename T>
inline void wall_push(const Consts<T>& c, const T* p, T dt, T* acc) {
    const T scale = c.wall_k / (dt * dt);
    for (int a = 0; a < 3; a++) {
        acc[a] = 0;
        if (p[a] < c.wall_d) {
            acc[a] = (c.wall_d - p[a]) * scale;
        } else if (p[a] > c.size - c.wall_d) {
            acc[a] = -((c.wall_d - (c.size - p[a])) * scale);
        }
    }
}

// GLSL min/max: min(x,y) = y < x ? y : x, max(x,y) = x < y ? y : x.
template <typename T>
inline T gmin(T x, T y) {
    return y < x ? y : x;
}
template <typename T>
inline T gmax(T x, T y) {
    return x < y ? y : x;
}

// update.comp:105-113 evaluated once per step (all operands are uniforms).  Always
// fp32: the host of the CUDA path performs the very same evaluation.
inline bool mouse_ray_hits_box(const wco_params* p) {
    float t1[3], t2[3];
    for (int a = 0; a < 3; a++) {
        const float tmin = (0.0f - p->mouse_origin[a]) / p->mouse_dir[a];
        const float tmax = (p->size - p->mouse_origin[a]) / p->mouse_dir[a];
        t1[a] = gmin(tmin, tmax);
        t2[a] = gmax(tmin, tmax);
    }
    const float tnear = gmax(gmax(t1[0], t1[1]), t1[2]);
    const float tfar = gmin(gmin(t2[0], t2[1]), t2[2]);
    return !(tnear > tfar);  // update.comp:119: "if (x > y) return vec3(0)"
}

// update.comp:116-132 for one particle (ray known to hit the box).
template <typename T>
inline void mouse_force(const Consts<T>& c, const T* pos, T pressure, T* out) {
    out[0] = out[1] = out[2] = 0;
    T to[3];
    for (int a = 0; a < 3; a++) to[a] = pos[a] - c.mo[a];
    // cross(mouseRayDirection, toMouse)
    const T cx = c.md[1] * to[2] - to[1] * c.md[2];
    const T cy = c.md[2] * to[0] - to[2] * c.md[0];
    const T cz = c.md[0] * to[1] - to[0] * c.md[1];
    const T d = std::sqrt(cx * cx + cy * cy + cz * cz);
    if (d > c.h) return;
    T w[3];
    spiky(c, to, d + c.eps, w);
    for (int a = 0; a < 3; a++) out[a] = ((-c.m * pressure) * w[a]) * (T)0.00001f;
}

// Visit every j != i of the 27-cell neighbourhood with dist < h, in the reference's
// order (density.comp:95-124 == update.comp:149-189).  The predicate is ALWAYS the
// fp32 one, so the fp64 variants see the same neighbour set.
template <typename F>
inline void for_each_neighbour(const wco_particle* P, int i, const uint32_t* counts,
                               const uint32_t* offsets, float bin, float h, int G, F&& body) {
    const float* pi = P[i].position;
    const int cx = cell_coord(pi[0], bin, G);
    const int cy = cell_coord(pi[1], bin, G);
    const int cz = cell_coord(pi[2], bin, G);
    for (int b = 0; b < 27; b++) {
        const int nx = cx + NEIGHBORHOOD[b].x, ny = cy + NEIGHBORHOOD[b].y,
                  nz = cz + NEIGHBORHOOD[b].z;
        if (nx < 0 || ny < 0 || nz < 0 || nx >= G || ny >= G || nz >= G) continue;
        const uint32_t index = (uint32_t)nz * G * G + (uint32_t)ny * G + (uint32_t)nx;
        const uint32_t count = counts[index];
        const uint32_t offset = offsets[index];
        for (uint32_t l = 0; l < count; l++) {
            const uint32_t j = offset + l;
            if (j == (uint32_t)i) continue;
            const float rx = pi[0] - P[j].position[0];
            const float ry = pi[1] - P[j].position[1];
            const float rz = pi[2] - P[j].position[2];
            const float dist = sqrtf(dist2_f32(rx, ry, rz));
            if (dist >= h) continue;
            body(j, rx, ry, rz, dist);
        }
    }
}

// density.comp:81-137 for particle i.  T=float is the oracle; T=double the "truth".
template <typename T>
inline void density_one(const wco_particle* P, int i, const uint32_t* counts,
                        const uint32_t* offsets, const Consts<T>& c, float bin_f, float h_f,
                        T* rho_out, T* pres_out, uint32_t* ncount) {
    T density = c.m * poly6(c, (T)0);
    uint32_t nn = 0;
    for_each_neighbour(P, i, counts, offsets, bin_f, h_f, c.G,
                       [&](uint32_t j, float rx, float ry, float rz, float dist) {
                           T d;
                           if (sizeof(T) == sizeof(float)) {
                               d = (T)dist;
                           } else {
                               const T x = (T)P[i].position[0] - (T)P[j].position[0];
                               const T y = (T)P[i].position[1] - (T)P[j].position[1];
                               const T z = (T)P[i].position[2] - (T)P[j].position[2];
                               d = std::sqrt(x * x + y * y + z * z);
                               (void)rx, (void)ry, (void)rz;
                           }
                           density += c.m * poly6(c, d);
                           nn++;
                       });
    T pos[3] = {(T)P[i].position[0], (T)P[i].position[1], (T)P[i].position[2]};
    if (c.phys & WCO_PHYS_WALL_PARTICLES)
        *rho_out = density + wall_density_particles(c, pos);
    else
        *rho_out = density + wall_density(c, pos);  // density.comp:126 (Q4: stored WITH wall term)
    const T q = density / c.rho0;               // density.comp:133 (Q4: pressure WITHOUT it)
    *pres_out = c.P0 + c.k * (((q * q) * q) - (T)1);
    if (ncount) *ncount = nn;
}

// update.comp:134-232 for particle i.  rho/pres arrays hold the (T-precision) density
// pass results for every particle.
template <typename T, typename RhoAt, typename PresAt>
inline void update_one(const wco_particle* P, int i, const uint32_t* counts,
                       const uint32_t* offsets, const Consts<T>& c, float bin_f, float h_f,
                       bool mouse_hits, T dt, RhoAt rho_at, PresAt pres_at, T* force_out,
                       T* vel_out, T* pos_out, T* term_scale_out = nullptr) {
    const T rho_i = rho_at(i), pres_i = pres_at(i);
    T pos[3] = {(T)P[i].position[0], (T)P[i].position[1], (T)P[i].position[2]};
    T vel_i[3] = {(T)P[i].velocity[0], (T)P[i].velocity[1], (T)P[i].velocity[2]};
    T Fp[3] = {0, 0, 0}, Fv[3] = {0, 0, 0};
    T ext[3] = {c.g[0] * rho_i, c.g[1] * rho_i, c.g[2] * rho_i};  // update.comp:145 (Q8)
    // test aid, not reference math: per component, the sum of the MAGNITUDES of everything
    // that is added into F -- the scale rounding errors of the sum are proportional to
    T S[3] = {std::abs(ext[0]), std::abs(ext[1]), std::abs(ext[2])};
    T cfN[3] = {0, 0, 0}, cfNabs[3] = {0, 0, 0}, cfL = 0, cfLabs = 0;

    for_each_neighbour(
        P, i, counts, offsets, bin_f, h_f, c.G,
        [&](uint32_t j, float rx, float ry, float rz, float dist) {
            T r[3], d;
            if (sizeof(T) == sizeof(float)) {
                r[0] = (T)rx, r[1] = (T)ry, r[2] = (T)rz;
                d = (T)dist;
            } else {
                for (int a = 0; a < 3; a++) r[a] = pos[a] - (T)P[j].position[a];
                d = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            }
            const T rho_j = rho_at(j), pres_j = pres_at(j);
            // update.comp:178 (Q9: only positive pair pressure pushes)
            const T pr = (pres_i + pres_j) / ((T)2 * rho_j);
            if (pr > (T)0) {
                T w[3];
                spiky(c, r, d + c.eps, w);  // update.comp:181 (Q7)
                for (int a = 0; a < 3; a++) Fp[a] -= (c.m * pr) * w[a];
                if (term_scale_out)
                    for (int a = 0; a < 3; a++) S[a] += std::abs((c.m * pr) * w[a]);
            }
            // update.comp:186-187
            const T wv = wvis(c, d);
            for (int a = 0; a < 3; a++) {
                const T vd = (T)P[j].velocity[a] - vel_i[a];
                Fv[a] += (c.m * (vd / rho_j)) * wv;
            }
            if (term_scale_out)  // both velocities enter the difference with their own rounding
                for (int a = 0; a < 3; a++)
                    S[a] += c.mu * std::abs((c.m * ((std::abs((T)P[j].velocity[a]) +
                                                     std::abs(vel_i[a])) / rho_j)) * wv);
            // colour field of the surface tension (extended physics): the sums are kept
            // without the common factor -6 * poly6C (grad W_poly6 = -6 C t^2 r,
            // lap W_poly6 = -6 C t (3 h^2 - 7 r^2), t = h^2 - r^2)
            if (c.phys & WCO_PHYS_SURFACE_TENSION) {
                const T d2 = sizeof(T) == sizeof(float) ? (T)dist2_f32(rx, ry, rz) : d * d;
                if (d2 > (T)0) {
                    const T t = c.h * c.h - d2;
                    const T a = c.m / rho_j;
                    for (int k = 0; k < 3; k++) {
                        cfN[k] += (a * (t * t)) * r[k];
                        cfNabs[k] += std::abs((a * (t * t)) * r[k]);
                    }
                    const T l = (a * t) * ((T)3 * (c.h * c.h) - (T)7 * d2);
                    cfL += l;
                    cfLabs += std::abs(l);
                }
            }
        });

    // update.comp:191
    T mf[3] = {0, 0, 0}, wf[3];
    if (mouse_hits) mouse_force(c, pos, pres_i, mf);
    if (c.phys & WCO_PHYS_WALL_PARTICLES) {
        // as a force density, so that the acceleration F / (rho_i + eps) below is the push
        wall_push(c, pos, dt, wf);
        for (int a = 0; a < 3; a++) wf[a] = wf[a] * (rho_i + c.eps);
    } else {
        wall_forces(c, pos, wf);
    }
    T st[3] = {0, 0, 0}, st_scale[3] = {0, 0, 0};
    if (c.phys & WCO_PHYS_SURFACE_TENSION) {
        {   // the particle's own term of the colour field's Laplacian (its gradient term is zero)
            const T l = (c.m / rho_i) * ((c.h * c.h) * ((T)3 * (c.h * c.h)));
            cfL += l;
            cfLabs += std::abs(l);
        }
        const T gc = (T)-6 * c.poly6C;
        const T n[3] = {gc * cfN[0], gc * cfN[1], gc * cfN[2]};
        const T lap = gc * cfL;
        const T len = std::sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        if (len > c.n_min) {
            const T f = (-c.sigma * lap) / len;
            for (int a = 0; a < 3; a++) {
                st[a] = f * n[a];
                // first-order error scale of f * n[a] (test aid): errors of lap and of n
                st_scale[a] = std::abs(c.sigma * gc) *
                              (cfLabs * std::abs(n[a]) / len + std::abs(lap / gc) * cfNabs[a] / len * (T)2);
            }
        }
    }
    for (int a = 0; a < 3; a++) {
        ext[a] += (mf[a] + wf[a]) + st[a];
        S[a] += std::abs(mf[a]) + std::abs(wf[a]) + std::abs(st[a]) + st_scale[a];
        if (term_scale_out) term_scale_out[a] = S[a];
    }

    T F[3], v[3], x[3];
    for (int a = 0; a < 3; a++) {
        Fv[a] *= c.mu;                            // update.comp:194
        F[a] = (Fp[a] + Fv[a]) + ext[a];          // update.comp:195
        const T acc = F[a] / (rho_i + c.eps);     // update.comp:198
        T vv = vel_i[a] + acc * dt;               // update.comp:199 (Q10: per component)
        vv = gmin(gmax(vv, (T)-50), (T)50);
        v[a] = vv;
        x[a] = pos[a] + vv * dt;                  // update.comp:200
    }
    // update.comp:202-227
    const T damping = (T)0.3f, border = (T)0.001f;
    for (int a = 0; a < 3; a++) {
        if (x[a] < border) {
            v[a] *= -damping;
            x[a] = border;
        } else if (x[a] > c.size - border) {
            v[a] *= -damping;
            x[a] = c.size - border;
        }
    }
    for (int a = 0; a < 3; a++) {
        if (force_out) force_out[a] = F[a];
        vel_out[a] = v[a];
        pos_out[a] = x[a];
    }
}

inline int clamp_threads(int nthreads) {
#ifdef _OPENMP
    // The caller's explicit count wins over OMP_NUM_THREADS (torchrun exports 1 to its
    // workers); it is only capped by the processors this process may run on.
    if (nthreads < 1) nthreads = 1;
    const int mx = omp_get_num_procs();
    return nthreads > mx ? mx : nthreads;
#else
    (void)nthreads;
    return 1;
#endif
}

}  // namespace

extern "C" {

int32_t wco_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Fluid::Fluid defaults, src/core/Fluid.cpp:9-27.
void wco_default_params(wco_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->num_particles = 80000;
    p->grid_res = 21;
    p->size = 1.0f;
    p->particle_radius = 0.01f;
    p->viscosity_coefficient = 200.0f;
    p->stiffness = 100.0f;
    p->rest_density = 500.0f;
    p->rest_pressure = 0.0f;
    p->gravity[0] = 0.0f;
    p->gravity[1] = -1.0f * 900.0f;
    p->gravity[2] = 0.0f;
    p->time_scale = 0.012f;
    // A ray that misses the box [0,size]^3 (Q19): mouse force == 0.
    p->mouse_origin[0] = p->mouse_origin[1] = p->mouse_origin[2] = -10.0f;
    p->mouse_dir[0] = -1.0f;
    p->mouse_dir[1] = p->mouse_dir[2] = 0.0f;
    // extended physics: off; the values used when a flag is set (wc_default_physics has the same)
    p->physics_flags = 0;
    p->surface_tension = 50.0f;
    p->surface_threshold = 7.0f;
    p->wall_stiffness = 0.5f;
    p->wall_distance = 0.01f;
    p->wall_rest_density = 0.0f;
}

// Fluid::setup, src/core/Fluid.cpp:206-216.  The kernel constants are evaluated in
// double from the float kernel radius (glm::pow(float, int) -> double context) and
// cast to float.
void wco_derive(const wco_params* p, wco_derived* d) {
    d->num_bins = p->grid_res * p->grid_res * p->grid_res;
    d->bin_size = p->size / (float)p->grid_res;
    d->kernel_radius = p->particle_radius * 4.0f;
    d->particle_mass = p->particle_radius * 8.0f;
    const double h = (double)d->kernel_radius;
    const double pi = 3.14159265358979323846;
    d->poly6_const = (float)(315.0 / (64.0 * pi * std::pow(h, 9)));
    d->spiky_const = (float)(-45.0 / (pi * std::pow(h, 6)));
    d->visc_const = (float)(45.0 / (pi * std::pow(h, 6)));
}

void wco_cell_ids(const wco_particle* in, int32_t n, float bin_size, int32_t grid_res,
                  uint32_t* cell_ids) {
    for (int32_t i = 0; i < n; i++) cell_ids[i] = cell_index(in[i].position, bin_size, grid_res);
}

void wco_sort(const wco_particle* in, int32_t n, float bin_size, int32_t grid_res,
              uint32_t* cell_ids, uint32_t* counts, uint32_t* offsets, uint32_t* perm,
              wco_particle* out) {
    const size_t B = (size_t)grid_res * grid_res * grid_res;
    std::vector<uint32_t> ids_local, counts_local, offsets_local;
    if (!cell_ids) {
        ids_local.resize(n);
        cell_ids = ids_local.data();
    }
    if (!counts) {
        counts_local.resize(B);
        counts = counts_local.data();
    }
    if (!offsets) {
        offsets_local.resize(B);
        offsets = offsets_local.data();
    }
    // count.comp:25-36 (Sort.cpp:255-256)
    std::memset(counts, 0, B * sizeof(uint32_t));
    for (int32_t i = 0; i < n; i++) {
        cell_ids[i] = cell_index(in[i].position, bin_size, grid_res);
        counts[cell_ids[i]]++;
    }
    // linearScan.comp:15-28 (index order z,y,x == linear order)
    uint32_t prefix = 0;
    for (size_t b = 0; b < B; b++) {
        offsets[b] = prefix;
        prefix += counts[b];
    }
    // reorder.comp:32-45 / sort.comp:32-45 with particleID ascending: the serial
    // loop makes "localOffset" the stable rank (Q1).  counts re-accumulates to the
    // same histogram (Sort.cpp:263-264), so it is left untouched here.
    std::vector<uint32_t> cursor(offsets, offsets + B);
    for (int32_t i = 0; i < n; i++) {
        const uint32_t dst = cursor[cell_ids[i]]++;
        if (perm) perm[dst] = (uint32_t)i;
        if (out) out[dst] = in[i];
    }
}

void wco_density(wco_particle* sorted, int32_t n, const uint32_t* counts,
                 const uint32_t* offsets, const wco_params* p, uint32_t* neighbour_counts,
                 int32_t nthreads) {
    const Consts<float> c = make_consts<float>(p);
    std::vector<float> rho(n), pres(n);
    const int nt = clamp_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int32_t i = 0; i < n; i++) {
        density_one<float>(sorted, i, counts, offsets, c, c.bin, c.h, &rho[i], &pres[i],
                           neighbour_counts ? &neighbour_counts[i] : nullptr);
    }
    // The gather reads positions only, so writing after the loop == the shader's
    // in-place write (report.pdf section 3.2.2).
    for (int32_t i = 0; i < n; i++) {
        sorted[i].density = rho[i];
        sorted[i].pressure = pres[i];
    }
}

void wco_update(const wco_particle* in, wco_particle* out, int32_t n, const uint32_t* counts,
                const uint32_t* offsets, const wco_params* p, float dt, float* forces,
                int32_t nthreads) {
    const Consts<float> c = make_consts<float>(p);
    const bool hits = mouse_ray_hits_box(p);
    const int nt = clamp_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int32_t i = 0; i < n; i++) {
        float v[3], x[3];
        update_one<float>(
            in, i, counts, offsets, c, c.bin, c.h, hits, dt,
            [&](uint32_t j) { return in[j].density; }, [&](uint32_t j) { return in[j].pressure; },
            forces ? &forces[3 * (size_t)i] : nullptr, v, x);
        wco_particle q = in[i];
        for (int a = 0; a < 3; a++) {
            q.velocity[a] = v[a];
            q.position[a] = x[a];
        }
        out[i] = q;
    }
}

void wco_step(wco_particle* buf1, wco_particle* buf2, int32_t n, const wco_params* p,
              float frame_dt, uint32_t* counts, uint32_t* offsets, int32_t nthreads) {
    wco_derived d;
    wco_derive(p, &d);
    // Fluid.cpp:347-350
    wco_sort(buf1, n, d.bin_size, p->grid_res, nullptr, counts, offsets, nullptr, buf2);
    wco_density(buf2, n, counts, offsets, p, nullptr, nthreads);
    wco_update(buf2, buf1, n, counts, offsets, p, frame_dt * p->time_scale, nullptr, nthreads);
}

void wco_advect(wco_particle* particles, int32_t n, float size, float dt) {
    const float damping = 0.3f, border = 0.01f;  // advect.comp:34-35
    for (int32_t i = 0; i < n; i++) {
        wco_particle& q = particles[i];
        for (int a = 0; a < 3; a++) {
            float v = q.velocity[a];
            float x = q.position[a] + v * dt;
            if (x < border) {
                v *= -damping;
                x = border;
            } else if (x > size - border) {
                v *= -damping;
                x = size - border;
            }
            q.position[a] = x;  // advect.comp:58 stores position only
            (void)v;
        }
    }
}

void wco_density_f64(const wco_particle* sorted, int32_t n, const uint32_t* counts,
                     const uint32_t* offsets, const wco_params* p, double* density,
                     double* pressure, int32_t nthreads) {
    const Consts<double> c = make_consts<double>(p);
    const Consts<float> cf = make_consts<float>(p);
    const int nt = clamp_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int32_t i = 0; i < n; i++) {
        density_one<double>(sorted, i, counts, offsets, c, cf.bin, cf.h, &density[i],
                            &pressure[i], nullptr);
    }
}

void wco_update_f64(const wco_particle* in, const double* density, const double* pressure,
                    int32_t n, const uint32_t* counts, const uint32_t* offsets,
                    const wco_params* p, float dt, double* forces, double* vel, double* pos,
                    int32_t nthreads) {
    wco_update_f64_scaled(in, density, pressure, n, counts, offsets, p, dt, forces, vel, pos,
                          nullptr, nthreads);
}

void wco_update_f64_scaled(const wco_particle* in, const double* density, const double* pressure,
                           int32_t n, const uint32_t* counts, const uint32_t* offsets,
                           const wco_params* p, float dt, double* forces, double* vel, double* pos,
                           double* term_scale, int32_t nthreads) {
    const Consts<double> c = make_consts<double>(p);
    const Consts<float> cf = make_consts<float>(p);
    const bool hits = mouse_ray_hits_box(p);
    const int nt = clamp_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int32_t i = 0; i < n; i++) {
        update_one<double>(
            in, i, counts, offsets, c, cf.bin, cf.h, hits, (double)dt,
            [&](uint32_t j) { return density[j]; }, [&](uint32_t j) { return pressure[j]; },
            forces ? &forces[3 * (size_t)i] : nullptr, &vel[3 * (size_t)i], &pos[3 * (size_t)i],
            term_scale ? &term_scale[3 * (size_t)i] : nullptr);
    }
}

}  // extern "C"
