// wc_sph_v1.cuh -- first correct CUDA path for density.comp / update.comp.
//
// One thread per cell-sorted particle.  Because x is the fastest-varying cell axis
// (count.comp:33), the 27-cell stencil is nine contiguous slices of the sorted array:
// for each (dy,dz) row, cells x-1..x+1 are adjacent, so slice = [offsets[row+x0],
// offsets[row+x1+1]).  Candidates are read straight from global memory (L1/L2); the
// warp-cooperative shared-memory version lives in wc_sph_tile.cuh and is checked against
// this one.  Kept as the simple, obviously-correct device path.
#pragma once

#include "wc_common.cuh"

namespace wc {

// density.comp:53-55 on the squared distance (h2 - d2 instead of h*h - dist*dist).
__device__ __forceinline__ float poly6_t3(float h2, float d2) {
    const float t = h2 - d2;
    return (t * t) * t;
}

// density.comp:57-79 (quirks Q2/Q3 reproduced; see oracle wall_density).
__device__ __forceinline__ float wall_density(const SphConsts& c, float x, float y, float z) {
    const float hi = c.size - c.h;
    float d = 0.0f;
    if (x < c.h) d += c.m * (poly6_t3(c.h2, x * x) * c.poly6C);
    else if (x > hi) { const float r = c.size - x; d += c.m * (poly6_t3(c.h2, r * r) * c.poly6C); }
    if (y < c.h) d += c.m * (poly6_t3(c.h2, y * y) * c.poly6C);
    else if (y > hi) { const float r = c.size - y; d += c.m * (poly6_t3(c.h2, r * r) * c.poly6C); }
    if (z < c.h) d += c.m * (poly6_t3(c.h2, z * z) * c.poly6C);
    else if (y > hi) { const float r = c.size - z; d += c.m * (poly6_t3(c.h2, r * r) * c.poly6C); }  // sic (density.comp:74)
    return d * 4.0f;
}

// Extended physics, WC_PHYS_WALL_PARTICLES (not in the reference; oracle wall_weight): the
// density of the wall particles as a function of the distance s to the wall only -- Harada et
// al.'s "wall weight function", here the closed-form integral of W_poly6 over the half space
// behind the wall: wall_w * (128/315 - P(s/h)), P(u) = u - 4u^3/3 + 6u^5/5 - 4u^7/7 + u^9/9.
__device__ __forceinline__ float wall_weight(const SphConstsExt& c, float s) {
    const float u = __fdiv_rn(s, c.h), u2 = u * u;
    float P = 1.0f / 9.0f;
    P = fmaf(P, u2, -4.0f / 7.0f);
    P = fmaf(P, u2, 6.0f / 5.0f);
    P = fmaf(P, u2, -4.0f / 3.0f);
    P = fmaf(P, u2, 1.0f);
    P = P * u;
    return c.wall_w * (128.0f / 315.0f - P);
}
__device__ __forceinline__ float wall_density_particles(const SphConstsExt& c, float x, float y,
                                                        float z) {
    const float hi = c.size - c.h;
    float d = 0.0f;
    if (x < c.h) d += wall_weight(c, x); else if (x > hi) d += wall_weight(c, c.size - x);
    if (y < c.h) d += wall_weight(c, y); else if (y > hi) d += wall_weight(c, c.size - y);
    if (z < c.h) d += wall_weight(c, z); else if (z > hi) d += wall_weight(c, c.size - z);
    return d;
}
// ... and the push that undoes a penetration of the wall's rest distance: the acceleration
// wall_acc * (wall_d - s) along the inward normal (oracle wall_push).
__device__ __forceinline__ float wall_push(const SphConstsExt& c, float x) {
    if (x < c.wall_d) return (c.wall_d - x) * c.wall_acc;
    if (x > c.size - c.wall_d) return -((c.wall_d - (c.size - x)) * c.wall_acc);
    return 0.0f;
}

// density.comp:126-133: stored density carries the wall term, pressure does not (Q4).
// kExt: the instantiation that knows the extended physics (chosen by c.phys at run time).
template <bool kExt = false>
__device__ __forceinline__ void finish_density(const ConstsOf<kExt>& c, float sum_t3, float x,
                                               float y, float z, float* rho_store, float* pres) {
    const float rho = (c.m * c.poly6C) * sum_t3;
    bool done = false;
    if constexpr (kExt) {
        if (c.phys & kPhysWall) {
            *rho_store = rho + wall_density_particles(c, x, y, z);
            done = true;
        }
    }
    if (!done) *rho_store = rho + wall_density(c, x, y, z);
    const float q = __fdividef(rho, c.rho0);
    *pres = c.P0 + c.k * ((q * q) * q - 1.0f);
}

// update.comp:71-100 with r.length() == 3 (Q5/Q6); same op order as the oracle.
__device__ __forceinline__ void wall_forces(const SphConsts& c, float x, float y, float z,
                                            float* fx, float* fy, float* fz) {
    const float s = (c.h - 3.0f) * (c.h - 3.0f);
    const float hi = c.size - c.h;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    if (x < c.h) ax = (s * __fdiv_rn(0.0f - x, 3.0f)) * c.spikyC;
    else if (x > hi) ax = (s * __fdiv_rn(c.size - x, 3.0f)) * c.spikyC;
    if (y < c.h) ay = (s * __fdiv_rn(0.0f - y, 3.0f)) * c.spikyC;
    else if (y > hi) ay = (s * __fdiv_rn(c.size - y, 3.0f)) * c.spikyC;
    if (z < c.h) az = (s * __fdiv_rn(0.0f - z, 3.0f)) * c.spikyC;
    else if (z > hi) az = (s * __fdiv_rn(c.size - z, 3.0f)) * c.spikyC;
    *fx = ax * 0.01f, *fy = ay * 0.01f, *fz = az * 0.01f;
}

// update.comp:116-132 for one particle, ray already known to hit the box.
__device__ __forceinline__ void mouse_force(const SphConsts& c, float x, float y, float z,
                                            float pres, float* fx, float* fy, float* fz) {
    const float tx = x - c.mo[0], ty = y - c.mo[1], tz = z - c.mo[2];
    const float cx = c.md[1] * tz - ty * c.md[2];
    const float cy = c.md[2] * tx - tz * c.md[0];
    const float cz = c.md[0] * ty - tx * c.md[1];
    const float d = __fsqrt_rn(cx * cx + cy * cy + cz * cz);
    *fx = *fy = *fz = 0.0f;
    if (d > c.h) return;
    const float dd = d + 1e-16f;
    const float s = (c.h - dd) * (c.h - dd);
    const float k = (-c.m * pres);
    *fx = (k * ((s * __fdiv_rn(tx, dd)) * c.spikyC)) * 0.00001f;
    *fy = (k * ((s * __fdiv_rn(ty, dd)) * c.spikyC)) * 0.00001f;
    *fz = (k * ((s * __fdiv_rn(tz, dd)) * c.spikyC)) * 0.00001f;
}

// Colour-field sums of the surface tension (extended physics), WITHOUT the common factor
// -6 * poly6C * m: N = sum t^2 r / rho_j, L = sum t (3 h^2 - 7 r^2) / rho_j over 0 < |r| < h
// (the particle's own term of L is added by integrate()).
struct ColourField {
    float Nx = 0, Ny = 0, Nz = 0, L = 0;
    __device__ __forceinline__ void add(const SphConstsExt& c, float rx, float ry, float rz,
                                        float d2, float inv_rho_j) {
        const float t = c.h2 - d2;
        const float at = inv_rho_j * t, att = at * t;
        Nx = fmaf(att, rx, Nx), Ny = fmaf(att, ry, Ny), Nz = fmaf(att, rz, Nz);
        const float l = at * fmaf(-7.0f, d2, c.h2x3);
        L += (d2 > 0.0f) ? l : 0.0f;
    }
};

// update.comp:191-231: external forces, symplectic Euler, speed clamp, box reflection.
// (Fp, Fv) are the gathered sums; Fv not yet scaled by the viscosity coefficient.
template <bool kExt = false>
__device__ __forceinline__ void integrate(const ConstsOf<kExt>& c, float4 pr, float4 vp, float Fpx,
                                          float Fpy, float Fpz, float Fvx, float Fvy, float Fvz,
                                          float4* pos_out, float4* vel_out, float4* force_out,
                                          const ColourField& cf = ColourField()) {
    float ex = c.g[0] * pr.w, ey = c.g[1] * pr.w, ez = c.g[2] * pr.w;  // update.comp:145 (Q8)
    float mx = 0.0f, my = 0.0f, mz = 0.0f, wx, wy, wz;
    if (c.mouse_hits) mouse_force(c, pr.x, pr.y, pr.z, vp.w, &mx, &my, &mz);
    bool walls_done = false, tension_done = false;
    if constexpr (kExt) {
        if (c.phys & kPhysWall) {  // as a force density: F / (rho + eps) below is the push
            const float rho_e = pr.w + 1e-16f;
            wx = wall_push(c, pr.x) * rho_e, wy = wall_push(c, pr.y) * rho_e,
            wz = wall_push(c, pr.z) * rho_e;
            walls_done = true;
        }
    }
    if (!walls_done) wall_forces(c, pr.x, pr.y, pr.z, &wx, &wy, &wz);
    if constexpr (kExt) {
        if (c.phys & kPhysTension) {  // F = -sigma * lap(c) * n / |n| at the surface
            const float nx = c.grad_m * cf.Nx, ny = c.grad_m * cf.Ny, nz = c.grad_m * cf.Nz;
            // (+ the particle's own term of the Laplacian; its gradient term is zero)
            const float lap = c.grad_m * (cf.L + __frcp_rn(pr.w) * (c.h2 * c.h2x3));
            const float len = __fsqrt_rn((nx * nx + ny * ny) + nz * nz);
            float sx = 0.0f, sy = 0.0f, sz = 0.0f;
            if (len > c.n_min) {
                const float f = __fdiv_rn(-c.sigma * lap, len);
                sx = f * nx, sy = f * ny, sz = f * nz;
            }
            ex += (mx + wx) + sx, ey += (my + wy) + sy, ez += (mz + wz) + sz;
            tension_done = true;
        }
    }
    if (!tension_done) ex += mx + wx, ey += my + wy, ez += mz + wz;
    const float Fx = (Fpx + Fvx * c.mu) + ex;
    const float Fy = (Fpy + Fvy * c.mu) + ey;
    const float Fz = (Fpz + Fvz * c.mu) + ez;
    const float inv = __frcp_rn(pr.w + 1e-16f);
    float vx = fminf(fmaxf(vp.x + (Fx * inv) * c.dt, -50.0f), 50.0f);  // Q10
    float vy = fminf(fmaxf(vp.y + (Fy * inv) * c.dt, -50.0f), 50.0f);
    float vz = fminf(fmaxf(vp.z + (Fz * inv) * c.dt, -50.0f), 50.0f);
    float x = pr.x + vx * c.dt, y = pr.y + vy * c.dt, z = pr.z + vz * c.dt;
    const float damping = 0.3f, border = 0.001f, top = c.size - border;
    if (x < border) { vx *= -damping; x = border; } else if (x > top) { vx *= -damping; x = top; }
    if (y < border) { vy *= -damping; y = border; } else if (y > top) { vy *= -damping; y = top; }
    if (z < border) { vz *= -damping; z = border; } else if (z > top) { vz *= -damping; z = top; }
    *pos_out = make_float4(x, y, z, pr.w);
    *vel_out = make_float4(vx, vy, vz, vp.w);
    if (force_out) *force_out = make_float4(Fx, Fy, Fz, 0.0f);
}

// One accepted pair of update.comp:174-187 (j != i, d2 < T).  Accumulates the pressure
// force WITHOUT the common factor (-m * spikyC) and the viscosity force WITHOUT
// (m * viscC); the caller applies both once.
__device__ __forceinline__ void pair_force(const SphConsts& c, float rx, float ry, float rz,
                                           float d2, float pres_i, float4 vi, float rho_j,
                                           float4 vj, float& Fpx, float& Fpy, float& Fpz,
                                           float& Fvx, float& Fvy, float& Fvz) {
    const float inv_rho = __frcp_rn(rho_j);
    const float inv_d = rsqrtf(fmaxf(d2, 1e-32f));  // Q7: dist == 0 -> r/d contributes 0
    const float dist = d2 * inv_d;
    const float hd = c.h - dist;
    const float pr = (pres_i + vj.w) * (0.5f * inv_rho);  // update.comp:178
    if (pr > 0.0f) {                                       // Q9
        const float w = pr * (hd * hd) * inv_d;
        Fpx = fmaf(w, rx, Fpx), Fpy = fmaf(w, ry, Fpy), Fpz = fmaf(w, rz, Fpz);
    }
    const float wv = hd * inv_rho;                         // update.comp:186-187
    Fvx = fmaf(wv, vj.x - vi.x, Fvx), Fvy = fmaf(wv, vj.y - vi.y, Fvy),
    Fvz = fmaf(wv, vj.z - vi.z, Fvz);
}

// advect.comp:20-59 (dead in the reference: its dispatch is commented out, Fluid.cpp:351).
// Position-only Euler step with the 0.01 border; only the position is stored (advect.comp:58).
__global__ void __launch_bounds__(256)
k_advect(float4* __restrict__ pos_rho, const float4* __restrict__ vel_pres, int n, float size,
         float dt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos_rho[i];
    const float4 v = vel_pres[i];
    const float border = 0.01f, top = size - border;
    float x = __fadd_rn(p.x, __fmul_rn(v.x, dt));
    float y = __fadd_rn(p.y, __fmul_rn(v.y, dt));
    float z = __fadd_rn(p.z, __fmul_rn(v.z, dt));
    x = x < border ? border : (x > top ? top : x);
    y = y < border ? border : (y > top ? top : y);
    z = z < border ? border : (z > top ? top : z);
    pos_rho[i] = make_float4(x, y, z, p.w);
}

template <bool kDebug>
__global__ void __launch_bounds__(128)
k_density_v1(float4* pos_rho, float4* __restrict__ vel_pres, const uint32_t* __restrict__ offsets,
             SphConstsExt c, uint32_t* __restrict__ neighbour_counts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c.n) return;
    const int i = c.first + t;
    const float4 p = pos_rho[i];
    const int G = c.G;
    const int cx = cell_coord(p.x, c.bin, G), cy = cell_coord(p.y, c.bin, G),
              cz = cell_coord(p.z, c.bin, G) - c.zbase;
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, G - 1);
    float acc = 0.0f;
    uint32_t nn = 0;
    for (int dz = -1; dz <= 1; dz++) {
        const int z = cz + dz;
        if (z < 0 || z >= c.Gz) continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int y = cy + dy;
            if (y < 0 || y >= G) continue;
            const uint32_t row = ((uint32_t)z * G + (uint32_t)y) * G;
            const uint32_t beg = offsets[row + x0], end = offsets[row + x1 + 1];
            for (uint32_t j = beg; j < end; j++) {
                const float4 q = pos_rho[j];
                const float d2 = dist2(p.x - q.x, p.y - q.y, p.z - q.z);
                if (d2 < c.T) {  // density.comp:117 ("dist >= h: skip"), self included (d2 = 0)
                    acc += poly6_t3(c.h2, d2);
                    if (kDebug) nn++;
                }
            }
        }
    }
    float rho, pres;
    finish_density<true>(c, acc, p.x, p.y, p.z, &rho, &pres);
    // In place like density.comp:135; the gather only reads x,y,z, which do not change.
    reinterpret_cast<float*>(pos_rho)[4 * (size_t)i + 3] = rho;
    reinterpret_cast<float*>(vel_pres)[4 * (size_t)i + 3] = pres;
    if (kDebug) neighbour_counts[t] = nn - 1u;  // the self pair is not a neighbour (j != i)
}

template <bool kDebug>
__global__ void __launch_bounds__(128)
k_update_v1(const float4* __restrict__ pos_rho, const float4* __restrict__ vel_pres,
            const uint32_t* __restrict__ offsets, SphConstsExt c, float4* __restrict__ pos_out,
            float4* __restrict__ vel_out, float4* __restrict__ forces) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c.n) return;
    const int i = c.first + t;
    const float4 p = pos_rho[i];
    const float4 v = vel_pres[i];
    const int G = c.G;
    const int cx = cell_coord(p.x, c.bin, G), cy = cell_coord(p.y, c.bin, G),
              cz = cell_coord(p.z, c.bin, G) - c.zbase;
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, G - 1);
    float Fpx = 0, Fpy = 0, Fpz = 0, Fvx = 0, Fvy = 0, Fvz = 0;
    ColourField cf;
    const bool tension = (c.phys & kPhysTension) != 0u;
    for (int dz = -1; dz <= 1; dz++) {
        const int z = cz + dz;
        if (z < 0 || z >= c.Gz) continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int y = cy + dy;
            if (y < 0 || y >= G) continue;
            const uint32_t row = ((uint32_t)z * G + (uint32_t)y) * G;
            const uint32_t beg = offsets[row + x0], end = offsets[row + x1 + 1];
            for (uint32_t j = beg; j < end; j++) {
                const float4 q = pos_rho[j];
                const float rx = p.x - q.x, ry = p.y - q.y, rz = p.z - q.z;
                const float d2 = dist2(rx, ry, rz);
                if (d2 < c.T && j != (uint32_t)i) {
                    pair_force(c, rx, ry, rz, d2, v.w, v, q.w, vel_pres[j], Fpx, Fpy, Fpz, Fvx,
                               Fvy, Fvz);
                    if (tension) cf.add(c, rx, ry, rz, d2, __frcp_rn(q.w));
                }
            }
        }
    }
    const float kp = -(c.m * c.spikyC), kv = c.m * c.viscC;
    float4 po, vo, fo;
    integrate<true>(c, p, v, Fpx * kp, Fpy * kp, Fpz * kp, Fvx * kv, Fvy * kv, Fvz * kv, &po, &vo,
                    kDebug ? &fo : nullptr, cf);
    pos_out[t] = po;
    vel_out[t] = vo;
    if (kDebug) forces[t] = fo;
}

}  // namespace wc
