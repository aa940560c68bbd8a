// lscpm_device.cuh — device-side data layout and the per-photon step for the LSC-PM photon tracer.
//
// What is restated here (scalar FP32, plate-local frame) is the loop body of pvtrace's
// photon_tracer.follow as driven by reference miniplant/simulation_runner.py:38-39, specialised to the
// scene family reference miniplant/scene_creator.py:52-274 builds: one plate Box holding parallel,
// equally spaced capillary Cylinders (outer tube + optional coaxial inner cylinder) whose end caps are
// coplanar with the plate's +-y faces.  The generic algorithm (brute force over every node, container =
// nearest node hit exactly once, adjacent = second intersection) is kept as the CPU oracle in oracle/;
// this file reproduces its outcomes with a medium state machine instead:
//
//   AIR   -> only the plate can be hit (entry), or nothing (exit/miss)
//   PLATE -> plate faces (container==hit -> adjacent = world) or the first capillary along the ray
//   OUTER -> in the tube wall of capillary k: inner cylinder, outer cylinder, or the +-y caps
//   INNER -> in the reaction mixture of capillary k: inner cylinder or the +-y caps
//
// including the literal consequences of pvtrace's adjacency rule (DESIGN.md "literal quirks"):
//   * leaving a tube wall with a sibling capillary ahead takes n2 from the sibling (adjacent = intersections[1]);
//   * leaving the inner cylinder when the straight continuation reaches a cap first takes n2 from the plate;
//   * a segment that ends on a cap (coincident with the plate's +-y face; ties resolve in scene-graph order,
//     plate first) is traced with the plate as container: plate material for absorption, n1 = plate, n2 = tube.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lscpm {

constexpr int MAX_COMP = 4;
constexpr int MED_AIR = 0, MED_PLATE = 1, MED_OUTER = 2, MED_INNER = 3;
constexpr int COMP_ABSORBER = 0, COMP_LUMINOPHORE = 1, COMP_REACTOR = 2;
constexpr int LIGHT_DIRECT = 0, LIGHT_DIFFUSE = 1;
constexpr int EV_GENERATE = 0, EV_REFLECT = 1, EV_TRANSMIT = 2, EV_ABSORB = 3, EV_EMIT = 4, EV_EXIT = 5, EV_KILL = 6,
              EV_REACT = 7;
constexpr int BIN_KILL = 0, BIN_MISS = 1, BIN_ENTRY_REFLECT = 2;
constexpr int SURF_NONE = 0, SURF_BOX = 1, SURF_OUTER = 2, SURF_INNER = 3, SURF_CAP = 4;
constexpr int MAXSTEPS = 1000;            // pvtrace follow(maxsteps=1000)
constexpr float ALPHA_ZERO = 1e-8f;       // numpy.isclose(alpha, 0) in Material.penetration_depth
constexpr float KB_EV = 8.617333262145e-5f;
constexpr float HC_EV_NM = 1240.8f;
#define LSCPM_INF __int_as_float(0x7f800000)

struct DevComponent {
    int kind;
    int abs_table;   // -1: constant coefficient
    int ems_table;   // luminophores only
    float coef;
    float qy;
};

struct DevMedium {
    float n;
    int n_comp;
    DevComponent comp[MAX_COMP];
};

// Lowered scene, passed to every kernel by value (constant bank).
struct DevScene {
    float hx, hy, hz;          // plate half extents
    float cx0, pitch, zc;      // capillary k: axis along y through (cx0 + k*pitch, zc)
    float inv_pitch;
    int n_ch;
    float r_o, r_i, r_o2, r_i2;
    int has_inner;
    DevMedium med[4];          // indexed by MED_*; med[MED_AIR] carries the world's refractive index only
    const float* tab_x;        // concatenated knot tables (scene spectra)
    const float* tab_y;
    const float* tab_cdf;      // trapezoid CDF, normalised (used for emission tables)
    const int* tab_begin;      // [n_tables + 1]
    int n_bins;
    int exit_base_plate;       // + box face
    int abs_base_plate;        // + component
    const int* abs_base_outer; // [n_ch]
    const int* abs_base_inner; // [n_ch]
    float w2l[9];              // rotation world -> plate-local (row-major)
};

struct DevBatch {
    int light_kind;
    int spec_begin;            // -1: constant wavelength
    int spec_n;
    uint32_t batch_id;
    int anchor_on_top;         // anchors lie on the plate's top face: a ray heading down enters at its anchor
    float wavelength;
    float mask_half[2];
    float A[12];               // mask frame -> plate-local, row-major 3x4
    float dir[3];              // DIRECT: plate-local unit direction
    float back_off;
    float cos_tilt, sin_tilt;  // DIFFUSE rejection test
};

struct Photon {
    float xr, y, z;            // x is stored relative to the axis of capillary kc
    float dx, dy, dz;
    float wl;
    float a_plate, a_outer, a_inner;   // total attenuation of the three media at wl
    int kc;
    int medium;
    int steps;                 // follow-loop iterations so far
    int n_surface;             // surface events so far
    int emitted;
};

struct Hit {
    float t;
    int kind;                  // SURF_*
    int face;                  // box face for SURF_BOX / SURF_CAP
    int ch;                    // capillary entered (SURF_OUTER from the plate)
    int container;             // medium whose material governs the segment
    int adj;                   // medium on the far side of the interface (pvtrace's `adjacent`)
    float n1, n2;
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter = (photon lo, photon hi, batch id, call index), key = seed.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

struct Rng {
    uint2 key;
    uint32_t p_lo, p_hi, batch;
    uint32_t calls;
    __device__ __forceinline__ uint4 next() { return philox4x32_10(make_uint4(p_lo, p_hi, batch, calls++), key); }
};

// 24-bit uniforms: u in [0,1) and its complement 1-u in (0,1], both exact in FP32
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
__device__ __forceinline__ float one_minus_u01(uint32_t x) { return (float)(((~x) >> 8) + 1u) * 5.9604644775390625e-8f; }

// ------------------------------------------------------------------------------------------------
// numpy.interp semantics on a knot table: clamped; segment = last j with xp[j] <= v
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float interp(float v, const float* __restrict__ xp, const float* __restrict__ fp, int n) {
    if (v <= __ldg(xp)) return __ldg(fp);
    if (v >= __ldg(xp + n - 1)) return __ldg(fp + n - 1);
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(xp + mid) <= v) lo = mid; else hi = mid;
    }
    const float x0 = __ldg(xp + lo), x1 = __ldg(xp + lo + 1), f0 = __ldg(fp + lo), f1 = __ldg(fp + lo + 1);
    return fmaf((f1 - f0) / (x1 - x0), v - x0, f0);
}

__device__ __forceinline__ float component_alpha(const DevScene& S, const DevComponent& c, float wl) {
    if (c.abs_table < 0) return c.coef;
    const int lo = __ldg(S.tab_begin + c.abs_table), n = __ldg(S.tab_begin + c.abs_table + 1) - lo;
    return interp(wl, S.tab_x + lo, S.tab_y + lo, n);
}

__device__ __forceinline__ float medium_alpha(const DevScene& S, int medium, float wl) {
    float a = 0.f;
    const DevMedium& m = S.med[medium];
#pragma unroll
    for (int c = 0; c < MAX_COMP; ++c)
        if (c < m.n_comp) a += component_alpha(S, m.comp[c], wl);
    return a;
}

__device__ __forceinline__ void refresh_alphas(const DevScene& S, Photon& p) {
    p.a_plate = medium_alpha(S, MED_PLATE, p.wl);
    p.a_outer = medium_alpha(S, MED_OUTER, p.wl);
    p.a_inner = S.has_inner ? medium_alpha(S, MED_INNER, p.wl) : 0.f;
}

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float plane_t(float target_pos, float target_neg, float o, float d) {
    // distance along d to the slab face the ray is heading to
    return d > 0.f ? (target_pos - o) / d : (d < 0.f ? (target_neg - o) / d : LSCPM_INF);
}

__device__ __forceinline__ float abs_x(const DevScene& S, const Photon& p) { return fmaf((float)p.kc, S.pitch, S.cx0) + p.xr; }

// exit of the plate box from inside
__device__ __forceinline__ void box_exit(const DevScene& S, const Photon& p, float& t, int& face) {
    const float tx = plane_t(S.hx, -S.hx, abs_x(S, p), p.dx);
    const float ty = plane_t(S.hy, -S.hy, p.y, p.dy);
    const float tz = plane_t(S.hz, -S.hz, p.z, p.dz);
    t = tx; face = p.dx > 0.f ? 1 : 0;
    if (ty < t) { t = ty; face = p.dy > 0.f ? 3 : 2; }
    if (tz < t) { t = tz; face = p.dz > 0.f ? 5 : 4; }
    t = fmaxf(t, 0.f);
}

// First capillary (outer cylinder) entered by a ray that is outside all capillaries, within tmax.
// Capillaries are parallel to y and ordered in x with disjoint x-extents, so the first hit in x order is the nearest.
__device__ __forceinline__ bool capillary_search(const DevScene& S, const Photon& p, float tmax, float& t_hit, int& j_hit) {
    const float a = p.dx * p.dx + p.dz * p.dz;
    if (!(a > 1e-30f)) return false;
    const float oz = p.z - S.zc;
    const float xe = fmaf(p.dx, tmax, p.xr);   // end of the segment, relative to capillary kc
    int j0, j1, step;
    if (p.dx >= 0.f) {
        j0 = p.kc + (int)ceilf((p.xr - S.r_o) * S.inv_pitch);
        j1 = p.kc + (int)floorf((xe + S.r_o) * S.inv_pitch);
        j0 = max(j0, 0); j1 = min(j1, S.n_ch - 1); step = 1;
    } else {
        j0 = p.kc + (int)floorf((p.xr + S.r_o) * S.inv_pitch);
        j1 = p.kc + (int)ceilf((xe - S.r_o) * S.inv_pitch);
        j0 = min(j0, S.n_ch - 1); j1 = max(j1, 0); step = -1;
    }
    const float ar2 = a * S.r_o2;
    for (int j = j0; step > 0 ? j <= j1 : j >= j1; j += step) {
        const float ox = p.xr - (float)(j - p.kc) * S.pitch;
        const float b = ox * p.dx + oz * p.dz;
        if (b >= 0.f) continue;                       // moving away from this axis
        const float cross = ox * p.dz - oz * p.dx;
        const float disc = ar2 - cross * cross;      // = b^2 - a*c without the cancellation
        if (disc <= 0.f) continue;
        const float c = ox * ox + oz * oz - S.r_o2;
        const float t = c / (sqrtf(disc) - b);
        if (t >= tmax) return false;                  // later capillaries are farther still
        if (t > 0.f) { t_hit = t; j_hit = j; return true; }
    }
    return false;
}

// far root of a cylinder of squared radius r2 about the axis of capillary kc, for a ray starting inside it
__device__ __forceinline__ float cylinder_exit(float ox, float oz, float dx, float dz, float r2) {
    const float a = dx * dx + dz * dz;
    if (!(a > 1e-30f)) return LSCPM_INF;
    const float b = ox * dx + oz * dz;
    const float cross = ox * dz - oz * dx;
    const float s = sqrtf(fmaxf(fmaf(a, r2, -cross * cross), 0.f));
    const float c = ox * ox + oz * oz - r2;           // <= 0 inside
    const float t = b > 0.f ? -c / (b + s) : (s - b) / a;
    return fmaxf(t, 0.f);
}

// near root for a ray starting outside (or on) the cylinder; INF when it does not enter
__device__ __forceinline__ float cylinder_enter(float ox, float oz, float dx, float dz, float r2) {
    const float a = dx * dx + dz * dz;
    const float b = ox * dx + oz * dz;
    if (!(b < 0.f)) return LSCPM_INF;
    const float cross = ox * dz - oz * dx;
    const float disc = fmaf(a, r2, -cross * cross);
    if (!(disc > 0.f)) return LSCPM_INF;
    const float c = ox * ox + oz * oz - r2;
    const float t = c / (sqrtf(disc) - b);
    return t > 0.f ? t : LSCPM_INF;
}

// ------------------------------------------------------------------------------------------------
// next interface for a photon in PLATE / OUTER / INNER (pvtrace next_hit, literal outcomes)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Hit find_hit(const DevScene& S, const Photon& p) {
    Hit h;
    h.ch = p.kc;
    h.face = 0;
    if (p.medium == MED_PLATE) {
        float t_box; int face;
        box_exit(S, p, t_box, face);
        float t_c; int j;
        h.container = MED_PLATE;
        h.n1 = S.med[MED_PLATE].n;
        if (capillary_search(S, p, t_box, t_c, j)) {
            h.t = t_c; h.kind = SURF_OUTER; h.ch = j; h.adj = MED_OUTER;
        } else {
            h.t = t_box; h.kind = SURF_BOX; h.face = face; h.adj = MED_AIR;
        }
        h.n2 = S.med[h.adj].n;
        return h;
    }
    const float oz = p.z - S.zc;
    const float t_cap = plane_t(S.hy, -S.hy, p.y, p.dy);
    const int cap_face = p.dy > 0.f ? 3 : 2;
    if (p.medium == MED_OUTER) {
        const float t_out = cylinder_exit(p.xr, oz, p.dx, p.dz, S.r_o2);
        const float t_in = S.has_inner ? cylinder_enter(p.xr, oz, p.dx, p.dz, S.r_i2) : LSCPM_INF;
        if (t_in < t_out && t_in < t_cap) {
            h.t = t_in; h.kind = SURF_INNER; h.container = MED_OUTER; h.adj = MED_INNER;
            h.n1 = S.med[MED_OUTER].n; h.n2 = S.med[MED_INNER].n;
        } else if (t_cap <= t_out) {
            // cap coincides with the plate face; the plate sorts first and becomes container and hit node
            h.t = fmaxf(t_cap, 0.f); h.kind = SURF_CAP; h.face = cap_face; h.container = MED_PLATE; h.adj = MED_OUTER;
            h.n1 = S.med[MED_PLATE].n; h.n2 = S.med[MED_OUTER].n;
        } else {
            h.t = t_out; h.kind = SURF_OUTER; h.container = MED_OUTER;
            h.n1 = S.med[MED_OUTER].n;
            // adjacent = whatever the straight continuation meets next: a sibling capillary, else the plate
            Photon q = p;
            q.xr = fmaf(p.dx, t_out, p.xr); q.y = fmaf(p.dy, t_out, p.y); q.z = fmaf(p.dz, t_out, p.z);
            float t_box; int face; box_exit(S, q, t_box, face);
            float t_s; int j;
            // the capillary being left cannot be re-entered: on its surface heading outward, b >= 0
            const bool sibling = capillary_search(S, q, t_box, t_s, j) && j != p.kc;
            h.adj = sibling ? MED_OUTER : MED_PLATE;
            if (sibling) h.ch = j;    // which sibling (only read by the probe; a leaving photon keeps kc)
            h.n2 = S.med[h.adj].n;
        }
        return h;
    }
    // MED_INNER
    const float t_side = cylinder_exit(p.xr, oz, p.dx, p.dz, S.r_i2);
    if (t_cap <= t_side) {
        h.t = fmaxf(t_cap, 0.f); h.kind = SURF_CAP; h.face = cap_face; h.container = MED_PLATE; h.adj = MED_OUTER;
        h.n1 = S.med[MED_PLATE].n; h.n2 = S.med[MED_OUTER].n;
    } else {
        h.t = t_side; h.kind = SURF_INNER; h.container = MED_INNER;
        h.n1 = S.med[MED_INNER].n;
        // adjacent = second intersection: the tube wall's outer surface, unless the cap (plate first) comes earlier
        const float t_wall = cylinder_exit(p.xr, oz, p.dx, p.dz, S.r_o2);
        h.adj = (t_cap <= t_wall) ? MED_PLATE : MED_OUTER;
        h.n2 = S.med[h.adj].n;
    }
    return h;
}

// entry of the plate box for a ray starting outside it (slab test); false when it misses
__device__ __forceinline__ bool box_enter(const DevScene& S, float ox, float oy, float oz, float dx, float dy, float dz,
                                          float& t, int& face) {
    float tmin = -LSCPM_INF, tmax = LSCPM_INF;
    int fmin = -1;
    const float h[3] = {S.hx, S.hy, S.hz}, o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (d[ax] == 0.f) {
            if (fabsf(o[ax]) > h[ax]) return false;
            continue;
        }
        float tlo = (-h[ax] - o[ax]) / d[ax], thi = (h[ax] - o[ax]) / d[ax];
        int flo = 2 * ax;
        if (tlo > thi) { const float tt = tlo; tlo = thi; thi = tt; flo = 2 * ax + 1; }
        if (tlo > tmin) { tmin = tlo; fmin = flo; }
        tmax = fminf(tmax, thi);
    }
    if (tmin > tmax || tmin <= 0.f) return false;
    t = tmin; face = fmin;
    return true;
}

// ------------------------------------------------------------------------------------------------
// optics: unpolarised Fresnel reflectivity with TIR, mirror reflection, Snell refraction
// (pvtrace FresnelSurfaceDelegate; reference scene uses the default surface everywhere)
// ------------------------------------------------------------------------------------------------
struct Interface {
    float nx, ny, nz;   // unit normal facing along the ray
    float cosi;         // d . n  (> 0)
    float k;            // cos of the refracted angle
    float R;
};

__device__ __forceinline__ Interface interface_eval(float dx, float dy, float dz, float nx, float ny, float nz,
                                                    float n1, float n2) {
    Interface I;
    float c = dx * nx + dy * ny + dz * nz;
    if (c < 0.f) { nx = -nx; ny = -ny; nz = -nz; c = -c; }
    c = fminf(c, 1.f);
    I.nx = nx; I.ny = ny; I.nz = nz; I.cosi = c;
    const float eta = n1 / n2;
    const float s2 = fmaxf(1.f - c * c, 0.f);
    const float k2 = 1.f - eta * eta * s2;
    if (n2 < n1 && k2 < 0.f) { I.R = 1.f; I.k = 0.f; return I; }   // beyond the critical angle
    const float k = sqrtf(fmaxf(k2, 0.f));
    const float rs = (n1 * c - n2 * k) / (n1 * c + n2 * k);
    const float rp = (n1 * k - n2 * c) / (n1 * k + n2 * c);
    I.k = k;
    I.R = 0.5f * (rs * rs + rp * rp);
    return I;
}

__device__ __forceinline__ void normalise(float& x, float& y, float& z) {
    const float s = rsqrtf(x * x + y * y + z * z);
    x *= s; y *= s; z *= s;
}

__device__ __forceinline__ void reflect_dir(Photon& p, const Interface& I) {
    const float f = 2.f * I.cosi;
    p.dx = fmaf(-f, I.nx, p.dx); p.dy = fmaf(-f, I.ny, p.dy); p.dz = fmaf(-f, I.nz, p.dz);
    normalise(p.dx, p.dy, p.dz);
}

__device__ __forceinline__ void refract_dir(Photon& p, const Interface& I, float n1, float n2) {
    const float eta = n1 / n2;
    const float g = I.k - eta * I.cosi;
    p.dx = fmaf(eta, p.dx, g * I.nx); p.dy = fmaf(eta, p.dy, g * I.ny); p.dz = fmaf(eta, p.dz, g * I.nz);
    normalise(p.dx, p.dy, p.dz);
}

__device__ __forceinline__ void box_normal(int face, float& nx, float& ny, float& nz) {
    nx = ny = nz = 0.f;
    const float s = (face & 1) ? 1.f : -1.f;
    if (face < 2) nx = s; else if (face < 4) ny = s; else nz = s;
}

// ------------------------------------------------------------------------------------------------
// photon generation (reference miniplant/utils.py:87-142, scene_creator.py:288-304)
// ------------------------------------------------------------------------------------------------
struct RayStart {
    float ox, oy, oz;   // plate-local origin of the ray (anchor - back_off * d)
    float ax, ay, az;   // anchor
    float dx, dy, dz;
    float wl;
};

__device__ __forceinline__ RayStart generate_ray(const DevScene& S, const DevBatch& B, const float* __restrict__ spec_x,
                                                 const float* __restrict__ spec_cdf, Rng& rng) {
    RayStart r;
    uint4 w = rng.next();
    const float mx = (2.f * u01(w.x) - 1.f) * B.mask_half[0];
    const float my = (2.f * u01(w.y) - 1.f) * B.mask_half[1];
    r.ax = fmaf(B.A[0], mx, fmaf(B.A[1], my, B.A[3]));
    r.ay = fmaf(B.A[4], mx, fmaf(B.A[5], my, B.A[7]));
    r.az = fmaf(B.A[8], mx, fmaf(B.A[9], my, B.A[11]));
    r.wl = B.spec_begin < 0 ? B.wavelength : interp(u01(w.z), spec_cdf + B.spec_begin, spec_x + B.spec_begin, B.spec_n);
    if (B.light_kind == LIGHT_DIRECT) {
        r.dx = B.dir[0]; r.dy = B.dir[1]; r.dz = B.dir[2];
    } else {
        float sz, cz, sa, ca;
        for (;;) {   // front-face rejection, uniform in (zenith, azimuth) angle
            const uint4 v = rng.next();
            sincospif(2.f * u01(v.x), &sa, &ca);      // azimuth = 360 deg * u
            sincospif(0.5f * u01(v.y), &sz, &cz);     // zenith  =  90 deg * u
            if (B.cos_tilt * cz + B.sin_tilt * sz * ca >= 0.f) break;
        }
        const float wx = -sz * ca, wy = -sz * sa, wz = -cz;   // world direction = -sky
        r.dx = S.w2l[0] * wx + S.w2l[1] * wy + S.w2l[2] * wz;
        r.dy = S.w2l[3] * wx + S.w2l[4] * wy + S.w2l[5] * wz;
        r.dz = S.w2l[6] * wx + S.w2l[7] * wy + S.w2l[8] * wz;
    }
    r.ox = fmaf(-B.back_off, r.dx, r.ax);
    r.oy = fmaf(-B.back_off, r.dy, r.ay);
    r.oz = fmaf(-B.back_off, r.dz, r.az);
    return r;
}

// which capillary (if any) contains the plate-local point; sets kc/xr/medium
__device__ __forceinline__ void locate_in_plate(const DevScene& S, Photon& p, float x_abs) {
    int k = (int)rintf((x_abs - S.cx0) * S.inv_pitch);
    k = min(max(k, 0), max(S.n_ch - 1, 0));
    p.kc = k;
    p.xr = x_abs - fmaf((float)k, S.pitch, S.cx0);
    p.medium = MED_PLATE;
    if (S.n_ch > 0) {
        const float oz = p.z - S.zc, rr = p.xr * p.xr + oz * oz;
        if (rr < S.r_o2) p.medium = (S.has_inner && rr < S.r_i2) ? MED_INNER : MED_OUTER;
    }
}

// keep xr relative to the nearest capillary axis while travelling through the plate
__device__ __forceinline__ void rebase_x(const DevScene& S, Photon& p) {
    const int shift = (int)rintf(p.xr * S.inv_pitch);
    int k = min(max(p.kc + shift, 0), max(S.n_ch - 1, 0));
    p.xr -= (float)(k - p.kc) * S.pitch;
    p.kc = k;
}

// ------------------------------------------------------------------------------------------------
// one iteration of follow() for a photon inside the plate assembly.
// Returns -1 while the photon lives, otherwise its tally bin.  `ev`/`ev_node_medium` report what happened.
// ------------------------------------------------------------------------------------------------
struct StepOut {
    int event;       // EV_*
    int medium;      // container medium of the event (for node lookup by recorders)
    int ch;          // capillary index of the event
    int kind;        // surface kind for surface events
    int face;
};

__device__ __forceinline__ int photon_step(const DevScene& S, Photon& p, Rng& rng, StepOut& out) {
    if (++p.steps > MAXSTEPS) { out.event = EV_KILL; return BIN_KILL; }
    const Hit h = find_hit(S, p);
    const uint4 w = rng.next();
    const float alpha = h.container == MED_PLATE ? p.a_plate : (h.container == MED_OUTER ? p.a_outer : p.a_inner);
    const float depth = alpha <= ALPHA_ZERO ? LSCPM_INF : -logf(one_minus_u01(w.x)) / alpha;
    out.medium = h.container; out.ch = p.kc; out.kind = h.kind; out.face = h.face;

    if (depth < h.t) {
        // absorbed inside the container: move, choose the component ~ alpha_i, radiative test
        p.xr = fmaf(p.dx, depth, p.xr); p.y = fmaf(p.dy, depth, p.y); p.z = fmaf(p.dz, depth, p.z);
        if (p.medium == MED_PLATE) rebase_x(S, p);
        const DevMedium& m = S.med[h.container];
        const float pick = u01(w.y) * alpha;
        float acc = 0.f;
        int index = m.n_comp - 1;
        bool found = false;
#pragma unroll
        for (int c = 0; c < MAX_COMP; ++c) {
            if (c < m.n_comp) {
                acc += component_alpha(S, m.comp[c], p.wl);
                if (!found && pick < acc) { index = c; found = true; }
            }
        }
        const DevComponent& comp = m.comp[index];
        if (comp.kind == COMP_LUMINOPHORE && u01(w.z) < comp.qy) {
            const uint4 v = rng.next();
            float sp, cp;
            sincospif(2.f * u01(v.x), &sp, &cp);
            const float mu = 2.f * u01(v.y) - 1.f;
            const float st = sqrtf(fmaxf(1.f - mu * mu, 0.f));
            p.dx = st * cp; p.dy = st * sp; p.dz = mu;
            const int lo = __ldg(S.tab_begin + comp.ems_table), n = __ldg(S.tab_begin + comp.ems_table + 1) - lo;
            const float ev = HC_EV_NM / p.wl + 1.5f * KB_EV * 300.f;
            const float p1 = interp(HC_EV_NM / ev, S.tab_x + lo, S.tab_cdf + lo, n);
            const float gamma = fmaf(1.f - p1, u01(v.z), p1);
            p.wl = interp(gamma, S.tab_cdf + lo, S.tab_x + lo, n);
            refresh_alphas(S, p);
            p.emitted = 1;
            out.event = EV_EMIT;
            return -1;
        }
        out.event = comp.kind == COMP_REACTOR ? EV_REACT : EV_ABSORB;
        const int base = h.container == MED_PLATE ? S.abs_base_plate
                       : (h.container == MED_OUTER ? __ldg(S.abs_base_outer + p.kc) : __ldg(S.abs_base_inner + p.kc));
        return base + index;
    }

    // reach the surface
    p.xr = fmaf(p.dx, h.t, p.xr); p.y = fmaf(p.dy, h.t, p.y); p.z = fmaf(p.dz, h.t, p.z);
    float nx, ny, nz;
    if (h.kind == SURF_BOX || h.kind == SURF_CAP) {
        box_normal(h.face, nx, ny, nz);
        // snap the coordinate of the face that was hit
        if (h.face < 2) { if (h.kind == SURF_BOX) { const float xa = (h.face & 1) ? S.hx : -S.hx; p.xr = xa - fmaf((float)p.kc, S.pitch, S.cx0); } }
        else if (h.face < 4) p.y = (h.face & 1) ? S.hy : -S.hy;
        else p.z = (h.face & 1) ? S.hz : -S.hz;
    } else {
        if (h.kind == SURF_OUTER && p.medium == MED_PLATE) {   // entering capillary h.ch from the plate
            p.xr -= (float)(h.ch - p.kc) * S.pitch;
            p.kc = h.ch;
        }
        const float oz = p.z - S.zc;
        const float inv = rsqrtf(p.xr * p.xr + oz * oz);
        nx = p.xr * inv; ny = 0.f; nz = oz * inv;
    }
    const Interface I = interface_eval(p.dx, p.dy, p.dz, nx, ny, nz, h.n1, h.n2);
    ++p.n_surface;
    out.ch = p.kc;
    if (u01(w.y) < I.R) {
        reflect_dir(p, I);
        out.event = EV_REFLECT;
        return -1;
    }
    refract_dir(p, I, h.n1, h.n2);
    out.event = EV_TRANSMIT;
    if (h.kind == SURF_BOX || h.kind == SURF_CAP) {
        // left the plate assembly: the next follow() iteration meets only the world sphere -> EXIT
        p.medium = MED_AIR;
        return (p.steps + 1 > MAXSTEPS) ? BIN_KILL : S.exit_base_plate + h.face;
    }
    if (h.kind == SURF_OUTER) p.medium = (p.medium == MED_PLATE) ? MED_OUTER : MED_PLATE;
    else p.medium = (p.medium == MED_OUTER) ? MED_INNER : MED_OUTER;
    if (p.medium == MED_PLATE) rebase_x(S, p);
    return -1;
}

// Start of a photon history.  A ray whose origin lies inside the plate starts there without an entry
// event (pvtrace finds the container geometrically); otherwise the first follow() iteration hits the
// plate (or misses it) and does the Fresnel test at the entry face.
// Returns -1 when the photon is alive inside the plate assembly, else its tally bin.
__device__ __forceinline__ int photon_start(const DevScene& S, Photon& p, float ox, float oy, float oz,
                                            bool have_anchor_entry, float ax, float ay, Rng& rng, StepOut& out) {
    p.steps = 0; p.n_surface = 0; p.emitted = 0;
    out.event = EV_GENERATE; out.medium = MED_AIR; out.kind = SURF_NONE; out.face = 0; out.ch = 0;
    float t; int face;
    float px, py, pz;
    if (have_anchor_entry) { face = 5; px = ax; py = ay; pz = S.hz; }
    else {
        if (fabsf(ox) < S.hx && fabsf(oy) < S.hy && fabsf(oz) < S.hz) {
            p.y = oy; p.z = oz;
            locate_in_plate(S, p, ox);
            return -1;
        }
        if (!box_enter(S, ox, oy, oz, p.dx, p.dy, p.dz, t, face)) { p.steps = 1; out.event = EV_EXIT; return BIN_MISS; }
        px = fmaf(p.dx, t, ox); py = fmaf(p.dy, t, oy); pz = fmaf(p.dz, t, oz);
        if (face < 2) px = (face & 1) ? S.hx : -S.hx; else if (face < 4) py = (face & 1) ? S.hy : -S.hy; else pz = (face & 1) ? S.hz : -S.hz;
    }
    p.steps = 1;
    p.y = py; p.z = pz;
    locate_in_plate(S, p, px);
    const int inside_medium = p.medium;
    float nx, ny, nz;
    box_normal(face, nx, ny, nz);
    const float n1 = S.med[MED_AIR].n, n2 = S.med[MED_PLATE].n;
    const Interface I = interface_eval(p.dx, p.dy, p.dz, nx, ny, nz, n1, n2);
    const uint4 w = rng.next();
    p.n_surface = 1;
    out.medium = MED_AIR; out.kind = SURF_BOX; out.face = face; out.ch = p.kc;
    if (u01(w.y) < I.R) {
        reflect_dir(p, I);
        p.medium = MED_AIR;
        out.event = EV_REFLECT;
        return BIN_ENTRY_REFLECT;
    }
    refract_dir(p, I, n1, n2);
    p.medium = inside_medium;
    out.event = EV_TRANSMIT;
    return -1;
}

}  // namespace lscpm
