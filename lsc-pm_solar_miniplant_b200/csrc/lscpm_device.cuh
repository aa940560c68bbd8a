// lscpm_device.cuh — device-side data layout and the per-photon step for the LSC-PM photon tracer.
//
// What is restated here (scalar FP32, plate-local frame) is the loop body of pvtrace's
// photon_tracer.follow as driven by reference miniplant/simulation_runner.py:38-39, specialised to the
// scene family reference miniplant/scene_creator.py:52-274 builds: one plate Box holding parallel,
// equally spaced capillary Cylinders (outer tube + optional coaxial inner cylinder) whose end caps are
// coplanar with the plate's +-y faces.  The generic algorithm (brute force over every node, container =
// nearest node hit exactly once, adjacent = hit node or container of the remaining intersections) is kept as
// the CPU oracle in oracle/; this file reproduces its outcomes with a medium state machine:
//
//   AIR   -> a fresh photon sitting on the plate face it is about to enter (entry Fresnel test)
//   PLATE -> plate faces (container==hit -> adjacent = world) or the capillary of the current cell
//   OUTER -> in the tube wall of capillary k: inner cylinder, outer cylinder, or the +-y caps
//   INNER -> in the reaction mixture of capillary k: inner cylinder or the +-y caps
//
// with pvtrace's bookkeeping for every interface (DESIGN.md "container and adjacent"): container = the nearest node the ray
// leaves exactly once; adjacent = the hit node when the photon enters it, and the container of what lies beyond (the
// same rule applied to the remaining intersections) when it leaves its own container - in generic position always the
// enclosing node: plate for a tube wall, tube wall for the mixture, air for the plate, whatever else lies ahead.
// Coincident faces: the capillary caps coincide with the plate's +-y faces (and the side PV boxes share its edge faces).
// pvtrace sorts distances each node computed in its own frame, so float64 round-off orders such ties; here the order is
// drawn from a hash of the photon's position (tie_bits).  The first of the tied nodes is hit node and - being left
// exactly once - container of the whole segment, the second is adjacent: a photon that enters the mixture within a
// millimetre of a tube end has the plate or the tube wall as container two times in three and then leaves through the
// cap instead of reacting.  Probes keep scene-graph order (plate, tube, mixture), like the oracle without a generator.
//
// Divergence control (what the ncu profile of the first version asked for):
//   * ONE code path computes every candidate distance (three slabs, near/far roots of both cylinders about
//     the current capillary axis) for all media and selects by medium — no per-medium branches;
//   * in the plate the photon marches cell by cell (a cell = the x-interval owned by one capillary), so each
//     iteration tests exactly one capillary; crossing a cell boundary is a virtual, uncounted step (free paths
//     are memoryless, so re-sampling the Beer-Lambert depth per sub-segment leaves the statistics unchanged);
//     scenes with a luminophore in the plate go further (plate_flight): straight to the next real event through
//     the lattice of capillary images that unfolding the total internal reflections at the z faces produces;
//   * table lookups are O(1) through bucket indices instead of binary searches;
//   * photon generation and luminophore re-emission share ONE "source" block at the end of the iteration (new
//     direction + wavelength by inverse CDF + attenuation refresh); a fresh photon's entry Fresnel test is made
//     there as well (scenes without outside boxes), with the fourth word of the generation draw;
//   * rare paths (sky-diffuse rejection loop, generic plate entry, irregular tables, cap ties) are out of line
//     and exchange scalars only, so the photon state never leaves registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lscpm {

constexpr int MAX_COMP = 4;
constexpr int MAX_TABLES = 12;           // scene tables + one total-attenuation table per medium
constexpr int MED_AIR = 0, MED_PLATE = 1, MED_OUTER = 2, MED_INNER = 3, MED_OUT = 4;
constexpr int MAX_EXT = 6;                // boxes outside the plate (the reference's photovoltaic absorbers)
constexpr float EXT_TOL = 2e-7f;          // metres along a face normal: "distance is zero" for faces shared between the plate and an outside box
constexpr int COMP_ABSORBER = 0, COMP_LUMINOPHORE = 1, COMP_REACTOR = 2;
constexpr int LIGHT_DIRECT = 0, LIGHT_DIFFUSE = 1;
constexpr int EV_GENERATE = 0, EV_REFLECT = 1, EV_TRANSMIT = 2, EV_ABSORB = 3, EV_EMIT = 4, EV_EXIT = 5, EV_KILL = 6,
              EV_REACT = 7, EV_NONE = -1;
constexpr int BIN_KILL = 0, BIN_MISS = 1, BIN_ENTRY_REFLECT = 2;
constexpr int SURF_NONE = 0, SURF_BOX = 1, SURF_OUTER = 2, SURF_INNER = 3, SURF_CAP = 4, SURF_CELL = 5;
#ifndef LSCPM_FLIGHT_COLUMNS
#define LSCPM_FLIGHT_COLUMNS 2
#endif
constexpr int FLIGHT_COLUMNS = LSCPM_FLIGHT_COLUMNS;   // capillary columns a plate flight looks at before it stops at a cell boundary
constexpr int MAXSTEPS = 1000;            // pvtrace follow(maxsteps=1000)
constexpr float ALPHA_ZERO = 1e-8f;       // numpy.isclose(alpha, 0) in Material.penetration_depth
constexpr float KB_EV = 8.617333262145e-5f;
constexpr float HC_EV_NM = 1240.8f;
constexpr int INV_CDF_BUCKETS = 1024;     // guide-table resolution for inverse-CDF lookups of scene tables
constexpr int SPEC_BUCKETS = 32;          // guide-table resolution for the per-batch solar spectra
#define LSCPM_INF __int_as_float(0x7f800000)

// Bounds checks of every computed table / tally / cell index (compute-sanitizer is closed on this GPU pool):
// build with -DLSCPM_BOUNDS_CHECK and run the GPU tests; a violated check prints its location and traps the kernel.
#ifdef LSCPM_BOUNDS_CHECK
#include <cstdio>
#define LSCPM_CHECK(cond) do { if (!(cond)) { printf("LSCPM_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define LSCPM_CHECK(cond) do { } while (0)
#endif

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
    int single_tab;      // index of the only tabulated component (its alpha = total - const_sum), or -1
    int alpha_table;     // table of the TOTAL attenuation (constants folded in) on the union of the knots, or -1
    float const_sum;     // sum of the constant coefficients
    DevComponent comp[MAX_COMP];
};

// knot table with O(1) forward lookup (value at abscissa) and guided inverse-CDF lookup
struct DevTable {
    int begin, n;        // knots [begin, begin+n) of tab_x / tab_y / tab_cdf
    int uniform;         // knots are equally spaced (tables whose knots sit on a regular grid are stored resampled
                         // onto it — piecewise-linear interpolation is unchanged); lookup is then pure arithmetic
    float x0, inv_w;     // uniform: knot spacing 1/inv_w.  otherwise bucket b covers x0 + [b, b+1) / inv_w
    int fwd_begin, fwd_n;// forward bucket index (uint16 knot numbers) in tab_fwd
    int inv_begin;       // inverse guide table (INV_CDF_BUCKETS + 1 entries) in tab_inv, or -1
};

// Opaque box outside the plate, axis-aligned in the plate frame (reference miniplant/scene_creator.py:99-206: the bottom
// PV under an air gap, four side PVs sharing a face with the plate edges; n = 3.4, alpha = 1e10 m^-1)
struct DevExtBox {
    float cx, cy, cz, hx, hy, hz;   // centre and half extents, plate-local absolute coordinates
    float n;
    int abs_bin;                    // absorb/<node>/0
    int exit_base;                  // exit/<node>/<face>
};

// Lowered scene, passed to every kernel by value (constant bank).
struct DevScene {
    float hx, hy, hz;          // plate half extents
    float cx0, pitch, zc;      // capillary k: axis along y through (cx0 + k*pitch, zc)
    float inv_pitch;
    float cell_lo_first, cell_hi_last;   // x-extent (relative to the axis) of the first / last cell: out to the plate faces
    int n_ch;                  // number of cells (>= 1; a plate without capillaries is one cell with r_o2 < 0)
    int has_caps;              // capillaries present
    float r_o, r_i, r_o2, r_i2;
    int has_inner;
    DevMedium med[5];          // indexed by MED_*; med[MED_AIR] and med[MED_OUT] carry the world's refractive index only
    int n_ext;                 // outside boxes (0 for the standard scene: none of the MED_OUT code runs then)
    DevExtBox ext[MAX_EXT];
    unsigned ext_touch[6];     // per plate face: bit i set when outside box i has a face in that face's plane (shares it)
    DevTable tab[MAX_TABLES];
    const float* tab_x;        // concatenated knot tables (scene spectra)
    const float* tab_y;
    const float* tab_cdf;      // cumulative-sum CDF (pvtrace Distribution), normalised
    const unsigned short* tab_fwd;
    const unsigned short* tab_inv;
    int n_bins;
    int exit_base_plate;       // + box face
    int abs_base_plate;        // + component
    const int* abs_base_outer; // [n_ch]
    const int* abs_base_inner; // [n_ch]
    const int* exit_base_outer;// [n_ch] exit/<tube>/<side, -cap, +cap>: a photon can leave through a cap it ties with the plate's face
    const int* exit_base_inner;// [n_ch]
    int cap_slot_lo, cap_slot_hi;   // cylinder face slot (1 = -cap, 2 = +cap) of the cap in the plate's -y / +y face
    float w2l[9];              // rotation world -> plate-local (row-major)
#ifdef LSCPM_DIAG              // diagnostic build only (scripts/dye_bias.py): alternatives to the restated pvtrace rules
    int diag_emit;             // 0 kT window (the rule), 1 redshift (no thermal allowance), 2 full emission spectrum, 3 window of 3 kT
#endif
};

struct DevBatch {
    int light_kind;
    int spec_begin;            // -1: constant wavelength
    int spec_n;
    int spec_guide;            // offset of this spectrum's guide table (SPEC_BUCKETS + 1 entries)
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
    float xr, y, z;            // x relative to the axis of capillary (cell) kc, z relative to the axis height zc
    float dx, dy, dz;
    float wl;
    float a_plate, a_outer, a_inner;   // total attenuation of the three media at wl
    int kc;
    int medium;
    int steps;                 // follow-loop iterations so far
    int aux;                   // bits 0-2 entry face, 4-5 medium behind it (fresh photon); bit 8 has emitted; 16-23 emission
                               // table of a pending re-emission.  Scenes with outside boxes also track, for the EXIT
                               // classification: bits 9-10 surface events (saturating), bit 11 first one was a reflection
                               // in air, bits 12-14 node of the last surface event in air (0 plate, i+1 box i), 24-26 its face
};

struct Hit {
    float t;
    int kind;                  // SURF_*
    int face;                  // box face for SURF_BOX / SURF_CAP (bits 3..: exit-bin offset of a cap hit by the tube or the
                               // mixture, see cap_tie); direction (+1/-1) for SURF_CELL
    int ch;                    // capillary of the event
    int container;             // medium whose material governs the segment
    int adj;                   // medium on the far side of the interface (pvtrace's `adjacent`)
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

// The ten round keys (key + round * Weyl constants) are the same for every call of a launch: the host computes them once
// and the trace kernels read them from the kernel-parameter constant bank as immediate operands of the XORs, instead of
// bumping two key registers per round and lane (2 of the ~7 instructions of a round).  Same function, same bits.
struct RoundKeys { uint2 k[10]; };
__host__ __device__ __forceinline__ RoundKeys philox_round_keys(uint2 key) {
    RoundKeys r;
    for (int round = 0; round < 10; ++round) {
        r.k[round] = key;
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return r;
}
__device__ __forceinline__ uint4 philox4x32_10_rk(uint4 c, const RoundKeys& R) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ R.k[round].x, lo1, hi0 ^ c.w ^ R.k[round].y, lo0);
    }
    return c;
}

struct Rng {
    uint2 key;
    uint32_t p_lo, p_hi, batch;
    uint32_t calls;
    __device__ __forceinline__ uint4 next() { return philox4x32_10(make_uint4(p_lo, p_hi, batch, calls++), key); }
    __device__ __forceinline__ uint4 next(const RoundKeys& R) { return philox4x32_10_rk(make_uint4(p_lo, p_hi, batch, calls++), R); }
};

// 24-bit uniforms: u in [0,1) and its complement 1-u in (0,1], both exact in FP32
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
__device__ __forceinline__ float one_minus_u01(uint32_t x) { return (float)(((~x) >> 8) + 1u) * 5.9604644775390625e-8f; }

// single-instruction SFU approximations (MUFU.RCP / MUFU.SQRT, ~1 ulp): the geometry is statistical, and the
// per-ray parity test (1e-5 relative) passes with them
#ifdef LSCPM_PRECISE   // diagnostic build: IEEE division / square root / logarithm
__device__ __forceinline__ float frcp(float x) { return 1.f / x; }
__device__ __forceinline__ float fsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float flog(float x) { return logf(x); }
__device__ __forceinline__ float frsqrt(float x) { return rsqrtf(x); }
#else
__device__ __forceinline__ float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * frcp(b); }
__device__ __forceinline__ float flog(float x) { return __logf(x); }
__device__ __forceinline__ float frsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif

// ------------------------------------------------------------------------------------------------
// table lookups with numpy.interp semantics (clamped; segment = last j with xp[j] <= v)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lerp_knots(float v, const float* __restrict__ xp, const float* __restrict__ fp, int j) {
    const float x0 = __ldg(xp + j), x1 = __ldg(xp + j + 1), f0 = __ldg(fp + j), f1 = __ldg(fp + j + 1);
    const float den = x1 - x0;
    return den > 0.f ? fmaf(fdiv(f1 - f0, den), v - x0, f0) : f0;
}

// value of column `col` (tab_y or tab_cdf) at abscissa v — irregular tables: bucket index + short walk
__device__ __noinline__ float table_at_irregular(const DevTable T, const float* __restrict__ xp, const float* __restrict__ fp,
                                                 const unsigned short* __restrict__ fwd, float v) {
    if (v <= __ldg(xp)) return __ldg(fp);
    if (v >= __ldg(xp + T.n - 1)) return __ldg(fp + T.n - 1);
    int b = (int)((v - T.x0) * T.inv_w);
    b = min(max(b, 0), T.fwd_n - 1);
    int j = __ldg(fwd + b);
    LSCPM_CHECK(j >= 0 && j + 1 < T.n);
    while (j + 2 < T.n && v >= __ldg(xp + j + 1)) ++j;
    while (j > 0 && v < __ldg(xp + j)) --j;
    return lerp_knots(v, xp, fp, j);
}

__device__ __forceinline__ float table_at(const DevScene& S, int t, float v, const float* __restrict__ col) {
    const DevTable& T = S.tab[t];
    if (T.uniform) {
        float f = (v - T.x0) * T.inv_w;
        f = fminf(fmaxf(f, 0.f), (float)(T.n - 1));      // numpy.interp clamps outside the table
        const int j = min((int)f, T.n - 2);
        LSCPM_CHECK(j >= 0 && j + 1 < T.n);
        const float y0 = __ldg(col + T.begin + j), y1 = __ldg(col + T.begin + j + 1);
        return fmaf(f - (float)j, y1 - y0, y0);
    }
    return table_at_irregular(T, S.tab_x + T.begin, col + T.begin, S.tab_fwd + T.fwd_begin, v);
}

// inverse CDF through a guide table: bracket [idx[b], idx[b+1]+1] then a (usually empty) bisection
__device__ __noinline__ float inverse_cdf(float u, const float* __restrict__ cdf, const float* __restrict__ xs, int n,
                                             const unsigned short* __restrict__ guide, int buckets) {
    if (u <= __ldg(cdf)) return __ldg(xs);
    if (u >= __ldg(cdf + n - 1)) return __ldg(xs + n - 1);
    // clamp in the float domain: nvcc 12.9 / sm_100a folded `min((int)(u * G), G - 1)` with a constant-propagated G
    // into VIADDMNMX with a wrong immediate inside the out-of-line clone of this function (emit / follow kernels)
    const float fb = (float)buckets;
    const int b = (int)fminf(u * fb, fb - 1.f);
    LSCPM_CHECK(b >= 0 && b < buckets);
    int lo = __ldg(guide + b), hi = min((int)__ldg(guide + b + 1) + 1, n - 1);
    LSCPM_CHECK(lo >= 0 && lo < n - 1 && hi > lo && hi <= n - 1);
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) <= u) lo = mid; else hi = mid;
    }
    return lerp_knots(u, cdf, xs, lo);
}

__device__ __forceinline__ float component_alpha(const DevScene& S, const DevComponent& c, float wl) {
    return c.abs_table < 0 ? c.coef : table_at(S, c.abs_table, wl, S.tab_y);
}

__device__ __forceinline__ float medium_alpha(const DevScene& S, int medium, float wl) {
    const DevMedium& m = S.med[medium];
    return m.alpha_table < 0 ? m.const_sum : table_at(S, m.alpha_table, wl, S.tab_y);
}

__device__ __forceinline__ void refresh_alphas(const DevScene& S, Photon& p) {
    p.a_plate = medium_alpha(S, MED_PLATE, p.wl);
    p.a_outer = S.has_caps ? medium_alpha(S, MED_OUTER, p.wl) : 0.f;
    p.a_inner = S.has_inner ? medium_alpha(S, MED_INNER, p.wl) : 0.f;
}

// component of medium `m` that absorbs, chosen with probability alpha_c / alpha (pvtrace Material.component)
__device__ __noinline__ int pick_component_generic(const DevScene& S, int medium, float wl, float pick) {
    const DevMedium& m = S.med[medium];
    float acc = 0.f;
    for (int c = 0; c < m.n_comp - 1; ++c) {
        acc += component_alpha(S, m.comp[c], wl);
        if (pick < acc) return c;
    }
    return m.n_comp - 1;
}

__device__ __forceinline__ int pick_component(const DevScene& S, int medium, float wl, float alpha, float pick) {
    const DevMedium& m = S.med[medium];
    if (m.n_comp == 1) return 0;
    if (m.n_comp == 2 && (m.single_tab >= 0 || m.alpha_table < 0)) {
        // one constant + one tabulated component (or two constants): no lookup needed
        const float a0 = m.comp[0].abs_table < 0 ? m.comp[0].coef : alpha - m.const_sum;
        return pick < a0 ? 0 : 1;
    }
    return pick_component_generic(S, medium, wl, pick);
}

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float axis_x(const DevScene& S, int k) { return fmaf((float)k, S.pitch, S.cx0); }
__device__ __forceinline__ float abs_x(const DevScene& S, const Photon& p) { return axis_x(S, p.kc) + p.xr; }

// near and far roots of a cylinder of squared radius r2 about the current axis; a,b,cross,rr shared.
// near = INF unless the ray is approaching and actually crosses; far is the exit distance for a ray inside.
__device__ __forceinline__ void cylinder_roots(float a, float inv_a, float b, float cross2, float rr, float r2,
                                               float& t_near, float& t_far) {
    const float disc = fmaf(a, r2, -cross2);            // = b^2 - a*c without the cancellation
    const float s = fsqrt(fmaxf(disc, 0.f));
    const float c = rr - r2;
    const float q = s + fabsf(b);
    const float r1 = fdiv(c, q), r2_ = q * inv_a;
    // b < 0 (approaching the axis): near = c/q, far = q/a;   b >= 0: near = -q/a (behind), far = -c/q
    const float near_ = b < 0.f ? r1 : -r2_;
    const float far_ = b < 0.f ? r2_ : -r1;
    t_near = (disc > 0.f && b < 0.f && near_ > 0.f) ? near_ : LSCPM_INF;
    t_far = fmaxf(far_, 0.f);
}

// pseudo-random bits that stand for float64 round-off ordering the intersections of coincident faces
__device__ __forceinline__ unsigned tie_bits(const Photon& p) {
    unsigned h = (__float_as_uint(p.xr) * 2654435761u) ^ (__float_as_uint(p.z) * 2246822519u) ^ (__float_as_uint(p.y) * 3266489917u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return h;
}

// Container and adjacent (packed 3 bits each) of a segment in a tube wall or in the mixture whose straight line
// ends in the cap plane.  There the tube's (and the mixture's) only remaining intersection ties with the plate's face; the
// first of the tied nodes is the nearest node left exactly once - the container of the WHOLE segment, also when the inner
// cylinder is hit before the cap - and, if the segment ends on the cap, the hit node as well (so for a SURF_CAP event the
// hit node is the container), with the second as adjacent.
// `ties`: order drawn from tb (two-way tie {plate, tube}: bit 0; three-way {plate, tube, mixture}: bits 8.. mod 6);
// otherwise scene-graph order (plate, tube, mixture).
// Bits 6..: for a cap that is hit by the tube or the mixture, the offset that turns `exit_base_plate + yface` into that
// node's own exit bin (exit/<tube or mixture>/<cap>); it travels in the upper bits of Hit::face.
__device__ __noinline__ unsigned cap_tie(const DevScene& S, unsigned tb, bool ties, bool is_ou, bool ou_inner, bool ou_cap, bool in_cap,
                                         bool tube_exit_cap, int kc, int yface) {
    const bool plate_first = ties ? (tb & 1u) != 0u : true;
    int container, adj;
    if (is_ou) {
        container = plate_first ? MED_PLATE : MED_OUTER;
        adj = ou_inner ? MED_INNER : (ou_cap ? (plate_first ? MED_OUTER : MED_PLATE) : MED_PLATE);
    } else if (in_cap) {
        const unsigned perm = ties ? (tb >> 8) % 6u : 0u;          // (first, second): PT PX TP TX XP XT
        container = perm < 2u ? MED_PLATE : (perm < 4u ? MED_OUTER : MED_INNER);
        adj = (perm == 2u || perm == 4u) ? MED_PLATE : ((perm == 0u || perm == 5u) ? MED_OUTER : MED_INNER);
    } else {   // leaves the mixture through its side; beyond it the tube's exit ties with the plate's face
        container = MED_INNER;
        adj = (tube_exit_cap && plate_first) ? MED_PLATE : MED_OUTER;
    }
    unsigned off = 0u;
    if ((is_ou ? ou_cap : in_cap) && container != MED_PLATE) {
        const int base = container == MED_OUTER ? __ldg(S.exit_base_outer + kc) : __ldg(S.exit_base_inner + kc);
        const int delta = base + (yface == 3 ? S.cap_slot_hi : S.cap_slot_lo) - (S.exit_base_plate + yface);
        LSCPM_CHECK(delta > 0);
        off = (unsigned)max(delta, 0);
    }
    return (unsigned)container | ((unsigned)adj << 3) | (off << 6);
}

// Next interface for a photon in PLATE / OUTER / INNER (pvtrace next_hit, literal outcomes); also returns the
// virtual cell crossing (SURF_CELL) for plate flights.  One code path for all media.
// PLATE_TOO = false: the caller never asks for a photon in the plastic (plate flights handle those), so the plate's own
// candidates are left out.
// TIES = false (probes): coincident faces in scene-graph order instead of the hashed order.
template <bool PLATE_TOO = true, bool TIES = true>
__device__ __forceinline__ Hit compute_hit(const DevScene& S, const Photon& p) {
    Hit h;
    h.ch = p.kc;
    // slabs, branch-free: distance to the face the ray is heading to; a zero direction component gives +INF
    // (rcp(0) = +INF times a positive gap)
    const bool first = p.kc == 0, last = p.kc == S.n_ch - 1;
    const float half = 0.5f * S.pitch;
    const float x_hi = last ? S.cell_hi_last : half, x_lo = first ? S.cell_lo_first : -half;
    const float gap_y = copysignf(S.hy, p.dy) - p.y;
    const float gap_z = copysignf(S.hz, p.dz) - S.zc - p.z;
    const float gap_x = (p.dx >= 0.f ? x_hi : x_lo) - p.xr;
    const float ty = p.dy != 0.f ? gap_y * frcp(p.dy) : LSCPM_INF;
    const float tz = p.dz != 0.f ? gap_z * frcp(p.dz) : LSCPM_INF;
    const float tx = p.dx != 0.f ? gap_x * frcp(p.dx) : LSCPM_INF;
    // cylinders about the current axis
    const float a = fmaf(p.dx, p.dx, p.dz * p.dz);
    const bool has_a = a > 1e-30f;
    const float inv_a = has_a ? fdiv(1.f, a) : 0.f;
    const float b = fmaf(p.xr, p.dx, p.z * p.dz);
    const float cross = fmaf(p.xr, p.dz, -p.z * p.dx);
    const float cross2 = cross * cross;
    const float rr = fmaf(p.xr, p.xr, p.z * p.z);
    float o_near, o_far, i_near, i_far;
    cylinder_roots(a, inv_a, b, cross2, rr, S.r_o2, o_near, o_far);
    cylinder_roots(a, inv_a, b, cross2, rr, S.r_i2, i_near, i_far);
    if (!has_a) { o_near = i_near = LSCPM_INF; o_far = i_far = LSCPM_INF; }
    if (!S.has_inner) i_near = LSCPM_INF;

    const int yface = p.dy > 0.f ? 3 : 2;
    // PLATE: box faces / cell boundary / the cell's capillary
    float t_pl = tz; int k_pl = SURF_BOX, f_pl = p.dz > 0.f ? 5 : 4, adj_pl = MED_AIR;
    if (PLATE_TOO) {
        if (ty < t_pl) { t_pl = ty; f_pl = yface; }
        if (tx < t_pl) {
            t_pl = tx;
            const bool plate_face = p.dx > 0.f ? last : first;
            k_pl = plate_face ? SURF_BOX : SURF_CELL;
            f_pl = plate_face ? (p.dx > 0.f ? 1 : 0) : (p.dx > 0.f ? 1 : -1);
        }
        if (o_near < t_pl) { t_pl = o_near; k_pl = SURF_OUTER; f_pl = 0; adj_pl = MED_OUTER; }
    }
    // OUTER: inner cylinder, else cap, else the tube's own outer surface (adjacent = container of what lies beyond = plate)
    const bool ou_inner = i_near < o_far && i_near < ty;
    const bool tube_exit_cap = ty <= o_far;                  // the straight line leaves the TUBE through the cap plane
    const bool ou_cap = !ou_inner && tube_exit_cap;
    const float t_ou = ou_inner ? i_near : (ou_cap ? ty : o_far);
    const int k_ou = ou_inner ? SURF_INNER : (ou_cap ? SURF_CAP : SURF_OUTER);
    // INNER: cap, else the inner surface with the tube wall beyond
    const bool in_cap = ty <= i_far;
    const float t_in = in_cap ? ty : i_far;
    const int k_in = in_cap ? SURF_CAP : SURF_INNER;

    const bool is_pl = PLATE_TOO && p.medium == MED_PLATE, is_ou = p.medium == MED_OUTER;
    h.t = fmaxf(is_pl ? t_pl : (is_ou ? t_ou : t_in), 0.f);
    h.kind = is_pl ? k_pl : (is_ou ? k_ou : k_in);
    h.face = is_pl ? f_pl : ((h.kind == SURF_CAP) ? yface : 0);
    // generic position: the photon's own node is the nearest node left exactly once
    h.container = is_pl ? MED_PLATE : (is_ou ? MED_OUTER : MED_INNER);
    h.adj = is_pl ? adj_pl : (is_ou ? (ou_inner ? MED_INNER : MED_PLATE) : MED_OUTER);
    // the segment's straight line ends in the cap plane, where the faces of plate, tube (and mixture) coincide: rare
    if (!is_pl && (tube_exit_cap || (!is_ou && in_cap))) {
        const unsigned r = cap_tie(S, TIES ? tie_bits(p) : 0u, TIES, is_ou, ou_inner, ou_cap, in_cap, tube_exit_cap, p.kc, yface);
        h.container = r & 7u; h.adj = (r >> 3) & 7u;
        h.face |= (int)(r >> 6) << 3;
    }
    return h;
}

// ------------------------------------------------------------------------------------------------
// optics: unpolarised Fresnel reflectivity with TIR, mirror reflection, Snell refraction
// (pvtrace FresnelSurfaceDelegate; reference scene uses the default surface everywhere)
// ------------------------------------------------------------------------------------------------
struct Interface {
    float nx, ny, nz;   // unit normal facing along the ray
    float cosi;         // d . n  (> 0)
    float k;            // cos of the refracted angle
    float eta;          // n1 / n2
    float R;
};

__device__ __forceinline__ Interface interface_eval(float dx, float dy, float dz, float nx, float ny, float nz,
                                                    float n1, float n2) {
    Interface I;
    float c = dx * nx + dy * ny + dz * nz;
    if (c < 0.f) { nx = -nx; ny = -ny; nz = -nz; c = -c; }
    c = fminf(c, 1.f);
    I.nx = nx; I.ny = ny; I.nz = nz; I.cosi = c;
    const float eta = fdiv(n1, n2);
    I.eta = eta;
    const float s2 = fmaxf(fmaf(-c, c, 1.f), 0.f);
    const float k2 = fmaf(-eta * eta, s2, 1.f);
    if (n2 < n1 && k2 < 0.f) { I.R = 1.f; I.k = 0.f; return I; }   // beyond the critical angle
    const float k = fsqrt(fmaxf(k2, 0.f));
    const float rs = fdiv(n1 * c - n2 * k, n1 * c + n2 * k);
    const float rp = fdiv(n1 * k - n2 * c, n1 * k + n2 * c);
    I.k = k;
    I.R = 0.5f * (rs * rs + rp * rp);
    return I;
}

__device__ __forceinline__ void normalise(float& x, float& y, float& z) {
    const float s = frsqrt(x * x + y * y + z * z);
    x *= s; y *= s; z *= s;
}

__device__ __forceinline__ void reflect_dir(Photon& p, const Interface& I) {
    const float f = 2.f * I.cosi;
    p.dx = fmaf(-f, I.nx, p.dx); p.dy = fmaf(-f, I.ny, p.dy); p.dz = fmaf(-f, I.nz, p.dz);
    normalise(p.dx, p.dy, p.dz);
}

__device__ __forceinline__ void refract_dir(Photon& p, const Interface& I) {
    const float g = I.k - I.eta * I.cosi;
    p.dx = fmaf(I.eta, p.dx, g * I.nx); p.dy = fmaf(I.eta, p.dy, g * I.ny); p.dz = fmaf(I.eta, p.dz, g * I.nz);
    normalise(p.dx, p.dy, p.dz);
}

__device__ __forceinline__ void box_normal(int face, float& nx, float& ny, float& nz) {
    const float s = (face & 1) ? 1.f : -1.f;
    nx = face < 2 ? s : 0.f;
    ny = (face >= 2 && face < 4) ? s : 0.f;
    nz = face >= 4 ? s : 0.f;
}

// which capillary (if any) contains the plate-local point; sets kc/xr/z/medium
__device__ __forceinline__ void locate_in_plate(const DevScene& S, Photon& p, float x_abs, float z_abs) {
    int k = (int)rintf((x_abs - S.cx0) * S.inv_pitch);
    k = min(max(k, 0), S.n_ch - 1);
    p.kc = k;
    p.xr = x_abs - axis_x(S, k);
    p.z = z_abs - S.zc;
    p.medium = MED_PLATE;
    const float rr = fmaf(p.xr, p.xr, p.z * p.z);
    if (rr < S.r_o2) p.medium = (S.has_inner && rr < S.r_i2) ? MED_INNER : MED_OUTER;
}

// sky-diffuse direction (reference miniplant/utils.py:102-122): zenith ~ U(0,90) deg, azimuth ~ U(0,360) deg — uniform
// in angle, not in solid angle — redrawn until the direction sees the plate's front face.  Returns the WORLD
// direction of travel (= -sky) and the advanced Philox call counter.
struct SkyDraw { float x, y, z; uint32_t calls; };
__device__ __noinline__ SkyDraw sample_sky_direction(float cos_tilt, float sin_tilt, uint32_t k0, uint32_t k1, uint32_t p_lo,
                                                     uint32_t p_hi, uint32_t batch, uint32_t calls) {
    float sz, cz, sa, ca;
    for (;;) {
        const uint4 v = philox4x32_10(make_uint4(p_lo, p_hi, batch, calls++), make_uint2(k0, k1));
        sincospif(2.f * u01(v.x), &sa, &ca);      // azimuth = 360 deg * u
        sincospif(0.5f * u01(v.y), &sz, &cz);     // zenith  =  90 deg * u
        if (cos_tilt * cz + sin_tilt * sz * ca >= 0.f) break;
    }
    SkyDraw d; d.x = -sz * ca; d.y = -sz * sa; d.z = -cz; d.calls = calls;
    return d;
}

// generic plate entry for a ray that starts outside the box (scalars in, scalars out); t < 0: miss
struct Entry { float t; int face; };
__device__ __noinline__ Entry plate_entry(float hx, float hy, float hz, float ox, float oy, float oz, float dx, float dy, float dz) {
    float tmin = -LSCPM_INF, tmax = LSCPM_INF;
    int fmin = -1;
    const float h[3] = {hx, hy, hz}, o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
    Entry e; e.t = -1.f; e.face = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (d[ax] == 0.f) {
            if (fabsf(o[ax]) > h[ax]) return e;
            continue;
        }
        float tlo = (-h[ax] - o[ax]) / d[ax], thi = (h[ax] - o[ax]) / d[ax];
        int flo = 2 * ax;
        if (tlo > thi) { const float tt = tlo; tlo = thi; thi = tt; flo = 2 * ax + 1; }
        if (tlo > tmin) { tmin = tlo; fmin = flo; }
        tmax = fminf(tmax, thi);
    }
    if (tmin > tmax || tmin <= 0.f) return e;
    e.t = tmin; e.face = fmin;
    return e;
}

// ------------------------------------------------------------------------------------------------
// the "source" block: photon generation (reference miniplant/utils.py:87-142, scene_creator.py:288-304) and
// luminophore re-emission (pvtrace Luminophore.emit, method 'kT') both end in "wavelength by inverse CDF, then
// refresh the attenuation coefficients"; they share that tail and one Philox draw `v`.
// ------------------------------------------------------------------------------------------------
struct WavelengthDraw {
    const float* cdf; const float* xs; const unsigned short* guide;
    int n, buckets;
    float u;
    float constant;      // used when n == 0
};

// put a ray that starts at plate-local (ox,oy,oz) on the plate: inside -> its medium, outside -> AIR on the entry
// face (or BIN_MISS).  `anchored`: the entry point is already known to be (ax, ay) on the top face.
template <bool EXT>
__device__ __forceinline__ int place_photon(const DevScene& S, Photon& p, float ox, float oy, float oz, bool anchored,
                                            float ax, float ay) {
    p.steps = 0;
    if (EXT && !(fabsf(ox) < S.hx && fabsf(oy) < S.hy && fabsf(oz) < S.hz)) {
        // outside boxes may shadow the plate: travel through the air from the true origin
        p.kc = 0; p.xr = ox - axis_x(S, 0); p.y = oy; p.z = oz - S.zc; p.aux = 0; p.medium = MED_OUT;
        return -1;
    }
    int face; float px, py, pz;
    if (anchored) { face = 5; px = ax; py = ay; pz = S.hz; }
    else {
        if (fabsf(ox) < S.hx && fabsf(oy) < S.hy && fabsf(oz) < S.hz) {   // born inside the plate: no entry event
            p.y = oy; locate_in_plate(S, p, ox, oz); p.aux = 0;
            return -1;
        }
        const Entry e = plate_entry(S.hx, S.hy, S.hz, ox, oy, oz, p.dx, p.dy, p.dz);
        if (e.t < 0.f) { p.steps = 1; return BIN_MISS; }
        face = e.face;
        px = fmaf(p.dx, e.t, ox); py = fmaf(p.dy, e.t, oy); pz = fmaf(p.dz, e.t, oz);
        if (face < 2) px = (face & 1) ? S.hx : -S.hx; else if (face < 4) py = (face & 1) ? S.hy : -S.hy; else pz = (face & 1) ? S.hz : -S.hz;
    }
    p.y = py;
    locate_in_plate(S, p, px, pz);
    p.aux = face | (p.medium << 4);
    p.medium = MED_AIR;
    return -1;
}

// generation-specific half: anchor point, direction, placement; fills the wavelength draw.  Returns -1, BIN_MISS or
// (ENTER) BIN_ENTRY_REFLECT.
// ENTER (trace kernels, scenes without outside boxes): the photon's first follow() iteration - the Fresnel test on the
// plate face it arrives at - is carried out here with the fourth word of the generation draw, instead of costing the warp
// a whole iteration of its own; `events` is the number of interaction events that made (0 or 1).
template <bool EXT, bool ENTER>
__device__ __forceinline__ int source_fresh(const DevScene& S, const DevBatch& B, const float* __restrict__ spec_x,
                                            const float* __restrict__ spec_cdf, const unsigned short* __restrict__ spec_guide,
                                            Rng& rng, const uint4 v, Photon& p, WavelengthDraw& wd, float* origin, unsigned& events) {
    events = 0u;
    const float mx = (2.f * u01(v.x) - 1.f) * B.mask_half[0];
    const float my = (2.f * u01(v.y) - 1.f) * B.mask_half[1];
    const float ax = fmaf(B.A[0], mx, fmaf(B.A[1], my, B.A[3]));
    const float ay = fmaf(B.A[4], mx, fmaf(B.A[5], my, B.A[7]));
    const float az = fmaf(B.A[8], mx, fmaf(B.A[9], my, B.A[11]));
    wd.cdf = spec_cdf + max(B.spec_begin, 0); wd.xs = spec_x + max(B.spec_begin, 0); wd.guide = spec_guide + B.spec_guide;
    wd.n = B.spec_begin < 0 ? 0 : B.spec_n; wd.buckets = SPEC_BUCKETS; wd.u = u01(v.z); wd.constant = B.wavelength;
    if (B.light_kind == LIGHT_DIRECT) {
        p.dx = B.dir[0]; p.dy = B.dir[1]; p.dz = B.dir[2];
    } else {
        const SkyDraw d = sample_sky_direction(B.cos_tilt, B.sin_tilt, rng.key.x, rng.key.y, rng.p_lo, rng.p_hi, rng.batch, rng.calls);
        rng.calls = d.calls;
        p.dx = S.w2l[0] * d.x + S.w2l[1] * d.y + S.w2l[2] * d.z;
        p.dy = S.w2l[3] * d.x + S.w2l[4] * d.y + S.w2l[5] * d.z;
        p.dz = S.w2l[6] * d.x + S.w2l[7] * d.y + S.w2l[8] * d.z;
    }
    const float ox = fmaf(-B.back_off, p.dx, ax), oy = fmaf(-B.back_off, p.dy, ay), oz = fmaf(-B.back_off, p.dz, az);
    if (origin) { origin[0] = ox; origin[1] = oy; origin[2] = oz; }
    const int bin = place_photon<EXT>(S, p, ox, oy, oz, B.anchor_on_top && p.dz < 0.f, ax, ay);
    if (ENTER && !EXT && bin < 0 && p.medium == MED_AIR) {
        float nx, ny, nz;
        box_normal(p.aux & 7, nx, ny, nz);
        const Interface I = interface_eval(p.dx, p.dy, p.dz, nx, ny, nz, S.med[MED_AIR].n, S.med[MED_PLATE].n);
        const bool reflected = u01(v.w) < I.R;
        if (reflected) reflect_dir(p, I); else refract_dir(p, I);
        p.steps = 1;
        events = 1u;
        if (reflected) return BIN_ENTRY_REFLECT;   // bounced off at first contact, leaves the scene
        p.medium = (p.aux >> 4) & 3;
    }
    return bin;
}

// emission-specific half: isotropic direction, kT-restricted CDF window of the luminophore stored in p.aux
__device__ __forceinline__ void source_emit(const DevScene& S, const uint4 v, Photon& p, WavelengthDraw& wd) {
    float sp, cp;
    __sincosf(fmaf(6.2831853071795865f, u01(v.x), -3.1415926535897932f), &sp, &cp);   // phi - pi in [-pi, pi): SFU accuracy ~1e-6
    const float mu = 2.f * u01(v.y) - 1.f;
    const float st = fsqrt(fmaxf(fmaf(-mu, mu, 1.f), 0.f));
    p.dx = st * cp; p.dy = st * sp; p.dz = mu;
    const int t = (p.aux >> 16) & 0xff;
    const DevTable& T = S.tab[t];
#ifdef LSCPM_DIAG
    const float ev = fdiv(HC_EV_NM, p.wl) + (S.diag_emit == 1 ? 0.f : (S.diag_emit == 3 ? 3.f : 1.5f) * KB_EV * 300.f);
    const float p1 = S.diag_emit == 2 ? 0.f : table_at(S, t, fdiv(HC_EV_NM, ev), S.tab_cdf);
#else
    const float ev = fdiv(HC_EV_NM, p.wl) + 1.5f * KB_EV * 300.f;
    const float p1 = table_at(S, t, fdiv(HC_EV_NM, ev), S.tab_cdf);
#endif
    wd.cdf = S.tab_cdf + T.begin; wd.xs = S.tab_x + T.begin; wd.guide = S.tab_inv + T.inv_begin;
    wd.n = T.n; wd.buckets = INV_CDF_BUCKETS; wd.u = fmaf(1.f - p1, u01(v.z), p1); wd.constant = p.wl;
    p.aux = (p.aux & 0xffff) | 256;     // has emitted
}

__device__ __forceinline__ void source_wavelength(Photon& p, const WavelengthDraw& wd) {
    p.wl = wd.n == 0 ? wd.constant : inverse_cdf(wd.u, wd.cdf, wd.xs, wd.n, wd.guide, wd.buckets);
}
__device__ __forceinline__ void source_finish(const DevScene& S, Photon& p, const WavelengthDraw& wd) {
    source_wavelength(p, wd);
    refresh_alphas(S, p);
}

// ------------------------------------------------------------------------------------------------
// scenes with boxes outside the plate (photovoltaic absorbers): the air around the plate becomes a medium with
// geometry of its own.  All of this is out of line and only runs when S.n_ext > 0.
// ------------------------------------------------------------------------------------------------
// the ray's reciprocal direction is computed once per ray (RayInv) and shared by every box tested
struct RayInv { float ix, iy, iz; };
__device__ __forceinline__ RayInv ray_inverse(float dx, float dy, float dz) {
    RayInv r;
    r.ix = dx != 0.f ? frcp(dx) : 0.f; r.iy = dy != 0.f ? frcp(dy) : 0.f; r.iz = dz != 0.f ? frcp(dz) : 0.f;
    return r;
}

__device__ __forceinline__ void slab_interval(float cx, float cy, float cz, float hx, float hy, float hz, float ox, float oy,
                                              float oz, float dx, float dy, float dz, const RayInv& inv, float& tmin, float& tmax,
                                              int& fmin, int& fmax) {
    tmin = -LSCPM_INF; tmax = LSCPM_INF; fmin = 0; fmax = 0;
    const float c[3] = {cx, cy, cz}, h[3] = {hx, hy, hz}, o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz}, iv[3] = {inv.ix, inv.iy, inv.iz};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float rel = o[ax] - c[ax];
        if (d[ax] == 0.f) {
            if (fabsf(rel) > h[ax]) { tmin = LSCPM_INF; tmax = -LSCPM_INF; }
            continue;
        }
        const float ta = (-h[ax] - rel) * iv[ax], tb = (h[ax] - rel) * iv[ax];
        const bool fwd = d[ax] > 0.f;                    // entering through the low face, leaving through the high one
        const float tlo = fwd ? ta : tb, thi = fwd ? tb : ta;
        const int flo = 2 * ax + (fwd ? 0 : 1), fhi = 2 * ax + (fwd ? 1 : 0);
        if (tlo > tmin) { tmin = tlo; fmin = flo; }
        if (thi < tmax) { tmax = thi; fmax = fhi; }
    }
}

// spatial gap (metres, along the face normal) that the ray parameter t corresponds to for face f
__device__ __forceinline__ float normal_gap(float t, int f, float dx, float dy, float dz) {
    return t * fabsf(f < 2 ? dx : (f < 4 ? dy : dz));
}

// outside box whose entry face coincides with plate face `face` where the photon leaves through it (the side PV boxes share
// the plate's edge faces): its entry ties with the plate's exit
__device__ __noinline__ int ext_touching(const DevScene& S, int face, float ox, float oy, float oz, float dx, float dy, float dz) {
    const RayInv inv = ray_inverse(dx, dy, dz);
    unsigned todo = S.ext_touch[face];
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1u;
        const DevExtBox& b = S.ext[i];
        float tmin, tmax; int f, fx;
        slab_interval(b.cx, b.cy, b.cz, b.hx, b.hy, b.hz, ox, oy, oz, dx, dy, dz, inv, tmin, tmax, f, fx);
        if (tmin <= tmax && tmax > 0.f && (f ^ 1) == face && fabsf(normal_gap(tmin, f, dx, dy, dz)) <= EXT_TOL) return i;
    }
    return -1;
}

// Nearest surface for a ray in the air around the plate: node 0 = plate, i+1 = outside box i, -1 = none.
// `skip` is the convex body the photon has just left or bounced off (0 plate, i+1 box i, -1 none): it cannot be met
// again before another interaction, so it is excluded by logic rather than by a distance tolerance (a tolerance on the
// exit parameter would let photons that clip a box corner within that tolerance escape).
// Returns true when the ray starts on (within EXT_TOL of) a face of box `node` heading inwards.
__device__ __forceinline__ bool outside_hit(const DevScene& S, float x, float y, float z, float dx, float dy, float dz, int skip,
                                            float& t_best, int& node, int& face) {
    t_best = LSCPM_INF; node = -1; face = 0;
    const RayInv inv = ray_inverse(dx, dy, dz);
    for (int i = 0; i < S.n_ext; ++i) {
        if (i + 1 == skip) continue;
        const DevExtBox& b = S.ext[i];
        float tmin, tmax; int f, fx;
        slab_interval(b.cx, b.cy, b.cz, b.hx, b.hy, b.hz, x, y, z, dx, dy, dz, inv, tmin, tmax, f, fx);
        if (!(tmin <= tmax) || !(tmax > 0.f)) continue;
        // on its face (gap along the face normal within EXT_TOL: coincident faces differ by float rounding), heading in
        if (normal_gap(tmin, f, dx, dy, dz) <= EXT_TOL) { t_best = 0.f; node = i + 1; face = f; return true; }
        if (tmin < t_best) { t_best = tmin; node = i + 1; face = f; }
    }
    if (skip != 0) {
        float tmin, tmax; int f, fx;
        slab_interval(0.f, 0.f, 0.f, S.hx, S.hy, S.hz, x, y, z, dx, dy, dz, inv, tmin, tmax, f, fx);
        if (tmin <= tmax && tmax > 0.f && normal_gap(tmin, f, dx, dy, dz) > EXT_TOL && tmin < t_best) { t_best = tmin; node = 0; face = f; }
    }
    return false;
}

struct OutState { float x, y, z, dx, dy, dz; int aux, steps; };
struct OutResult { float x, y, z, dx, dy, dz; int aux, steps, bin, event, node, face, extra; bool entered; };

__device__ __forceinline__ int exit_bin(const DevScene& S, int aux) {
    const int n_surface = (aux >> 9) & 3;
    if (n_surface == 0) return BIN_MISS;
    if (n_surface == 1 && (aux & (1 << 11)) && !(aux & 256)) return BIN_ENTRY_REFLECT;
    const int node = (aux >> 12) & 7, face = (aux >> 24) & 7;
    if (node == 7) {   // left through a capillary end cap that was the hit node: cap node in bits 27-28, capillary in bits 16-23
        const int kc = (aux >> 16) & 0xff;
        const int base = ((aux >> 27) & 3) == MED_OUTER ? __ldg(S.exit_base_outer + kc) : __ldg(S.exit_base_inner + kc);
        return base + (face == 3 ? S.cap_slot_hi : S.cap_slot_lo);
    }
    return (node == 0 ? S.exit_base_plate : S.ext[node - 1].exit_base) + face;
}

__device__ __forceinline__ int note_surface(int aux, bool air_reflection, int node, int face, bool record) {
    const int n_surface = (aux >> 9) & 3;
    if (n_surface == 0 && air_reflection) aux |= 1 << 11;
    aux = (aux & ~(3 << 9)) | (min(n_surface + 1, 2) << 9);
    if (record) aux = (aux & ~((7 << 12) | (7 << 24))) | (node << 12) | (face << 24);
    return aux;
}

// the reflections a plate flight folded in: each one is a surface interaction inside the plate (nothing recorded)
__device__ __forceinline__ int note_folded(int aux, int bounces) {
    const int n_surface = (aux >> 9) & 3;
    return (aux & ~(3 << 9)) | (min(n_surface + bounces, 2) << 9);
}

// One follow() iteration of a photon travelling in the air around the plate: next hit among the plate and the outside
// boxes (container = world, adjacent = hit), Fresnel test; an outside box that is entered absorbs at once (alpha ~1e10).
__device__ __noinline__ OutResult outside_step(const DevScene& S, OutState q, float u_surf) {
    OutResult r;
    r.x = q.x; r.y = q.y; r.z = q.z; r.dx = q.dx; r.dy = q.dy; r.dz = q.dz; r.aux = q.aux; r.steps = q.steps + 1;
    r.bin = -1; r.event = EV_NONE; r.node = 0; r.face = 0; r.extra = 0; r.entered = false;
    if (r.steps > MAXSTEPS) { r.event = EV_KILL; r.bin = BIN_KILL; return r; }
    float t_best; int node, face;
    const int skip = ((q.aux >> 9) & 3) ? ((q.aux >> 12) & 7) : -1;
    if (outside_hit(S, q.x, q.y, q.z, q.dx, q.dy, q.dz, skip, t_best, node, face)) {
        // the photon sits on a face of box `node`, heading in: the zero-distance entry is dropped, it is inside the box
        r.event = EV_ABSORB; r.node = node; r.bin = S.ext[node - 1].abs_bin;
        return r;
    }
    if (node < 0) {   // nothing but the world sphere ahead
        r.event = EV_EXIT;
        r.bin = exit_bin(S, q.aux);
        return r;
    }
    r.x = fmaf(q.dx, t_best, q.x); r.y = fmaf(q.dy, t_best, q.y); r.z = fmaf(q.dz, t_best, q.z);
    float nx, ny, nz;
    box_normal(face, nx, ny, nz);
    const float n1 = S.med[MED_AIR].n, n2 = node == 0 ? S.med[MED_PLATE].n : S.ext[node - 1].n;
    const Interface I = interface_eval(q.dx, q.dy, q.dz, nx, ny, nz, n1, n2);
    Photon tmp; tmp.dx = q.dx; tmp.dy = q.dy; tmp.dz = q.dz;
    r.node = node; r.face = face;
    if (u_surf < I.R) {
        reflect_dir(tmp, I);
        r.dx = tmp.dx; r.dy = tmp.dy; r.dz = tmp.dz;
        r.event = EV_REFLECT;
        r.aux = note_surface(q.aux, true, node, face, true);
        return r;
    }
    refract_dir(tmp, I);
    r.dx = tmp.dx; r.dy = tmp.dy; r.dz = tmp.dz;
    r.event = EV_TRANSMIT;
    r.aux = note_surface(q.aux, false, node, face, false);
    if (node == 0) {
        // snap onto the plate face; the caller locates the medium behind it
        if (face < 2) r.x = (face & 1) ? S.hx : -S.hx; else if (face < 4) r.y = (face & 1) ? S.hy : -S.hy; else r.z = (face & 1) ? S.hz : -S.hz;
        r.entered = true;
        return r;
    }
    // transmitted into an opaque box: absorbed within nanometres (a second follow() iteration: ABSORB)
    if (r.steps + 1 > MAXSTEPS) { r.event = EV_KILL; r.bin = BIN_KILL; return r; }
    r.steps += 1; r.extra = 1;
    r.bin = S.ext[node - 1].abs_bin;
    return r;
}

// ------------------------------------------------------------------------------------------------
// one iteration of follow() with the iteration's Philox draw `w` (x: free path, y: Fresnel / component pick,
// z: quantum yield).  Returns -1 while the photon lives, STEP_EMIT when it was absorbed by a luminophore that
// re-emits (the source block then gives it a new direction and wavelength), otherwise its tally bin.
// ------------------------------------------------------------------------------------------------
constexpr int STEP_EMIT = -2;

struct StepOut {
    int event;       // EV_* (EV_NONE for a virtual cell crossing)
    int medium;      // container medium of the event (for node lookup by recorders)
    int ch;          // capillary index of the event
    int kind;        // surface kind for surface events
    int face;
    int extra;       // additional interaction events folded into this iteration (photon absorbed by an outside box)
};

// ------------------------------------------------------------------------------------------------
// The iteration is written in two halves, cut where photons stop doing the same thing: `advance` (nearest interface,
// free path, move: identical work for every photon in flight) ends by naming the interaction the photon is waiting for,
// the `interact_*` functions carry it out.  photon_step (below) runs them back to back for the lane-per-photon kernel
// and the recorders; the block-sorted kernel re-deals the block's photons between the two halves.
// ------------------------------------------------------------------------------------------------
// class order = deal order: FLAT and CYL share the Fresnel code, CELL has no work of its own, ABSORB and FRESH share the source code
constexpr int CLS_FLAT = 0, CLS_CYL = 1, CLS_CELL = 2, CLS_ABSORB = 3, CLS_FRESH = 4, CLS_OUT = 5, CLS_IDLE = 6, N_CLS = 7;

// pending-interaction word: class | surface kind | container | adjacent | face + 1 (the rest of the word)
__device__ __forceinline__ unsigned pack_pending(int cls, int kind, int face, int container, int adj) {
    return (unsigned)cls | ((unsigned)kind << 4) | ((unsigned)container << 7) | ((unsigned)adj << 10) | ((unsigned)(face + 1) << 13);
}
__device__ __forceinline__ int pending_kind(unsigned w) { return (w >> 4) & 7; }
__device__ __forceinline__ int pending_container(unsigned w) { return (w >> 7) & 7; }
__device__ __forceinline__ int pending_adj(unsigned w) { return (w >> 10) & 7; }
__device__ __forceinline__ int pending_face(unsigned w) { return (int)(w >> 13) - 1; }

// ------------------------------------------------------------------------------------------------
// Plate flight (standard scene, photon in the PMMA): straight to the next REAL event.
// pvtrace follows a photon that is trapped by total internal reflection bounce by bounce, and this kernel used to add a
// virtual step at every cell boundary; both are pure geometry.  Unfold the plate instead: mirror it at its z faces and
// the trapped photon travels a straight line through a lattice of capillary images - columns at x = cx0 + k*pitch,
// images of the axis at z' = m*T for even m and m*T - 2*zc for odd m (T = plate thickness).  One free path L is drawn for
// the whole flight (free paths are memoryless, so this is the same distribution as one draw per segment), the flight ends
// at the first of: absorption at L, a y or x face of the plate, a z face if the photon is NOT trapped (Fresnel test with
// R < 1: a real random event), the first capillary image hit.  Then the end point is folded back: slab index m gives
// the number of reflections (each one a follow() iteration and an interaction event of the reference, so both counters
// advance by |m|), z and dz are mirrored for odd m.  Trapped is decided by interface_eval's own expression.
// FOLD = false (path recorders, probe): the z faces always end the flight, so every reflection stays a step of its own.
// ------------------------------------------------------------------------------------------------
// Every product and sum below is an explicit intrinsic: the two trace kernels inline this code separately and must
// round identically (a fused multiply-add contracted in one copy only moves a grazing hit in 2e-4 of the photons and
// breaks the bit-identity the kernels are tested for).
//
// first intersection with the image circle centred (0, c) in the frame of its column; `cross0` is the cross product for
// c = 0, ra2 = a * r_o^2.  Returns INF when the line misses, the circle is behind, or c lies outside [c_lo, c_hi].
__device__ __forceinline__ float image_hit(float ox, float z, float dx, float dz, float cross0, float ra2, float r_o2, float c,
                                           float c_lo, float c_hi) {
    const float oz = __fsub_rn(z, c);
    const float b = fmaf(ox, dx, __fmul_rn(oz, dz));
    const float cr = fmaf(c, dx, cross0);
    const float disc = fmaf(-cr, cr, ra2);
    const float cc = __fsub_rn(fmaf(ox, ox, __fmul_rn(oz, oz)), r_o2);
    const float t = __fmul_rn(cc, frcp(__fsub_rn(fsqrt(fmaxf(disc, 0.f)), b)));
    const bool ok = c >= c_lo && c <= c_hi && disc > 0.f && b < 0.f && t > 0.f;
    return ok ? t : LSCPM_INF;
}

// images of column `ox` (x of the photon relative to the column's axis) met by an upward ray (dz >= 0): the image
// centres form two lattices of period 2T - even images at n*2T, odd (mirrored) ones at T + delta + n*2T.  Candidates: the
// first point of either lattice inside the interval of centres whose circle the line crosses, going up; in the photon's
// own column (`two`) that image can lie behind it, so the next one in order is tried as well.  Branch-free.
struct ColumnHit { float t, c; int odd; };
__device__ __forceinline__ ColumnHit column_images(float ox, float z, float dx, float dz, float ra, float ra2, float r_o, float r_o2,
                                                   float inv_dx, float T, float delta, bool two) {
    const float two_T = 2.f * T, inv_2T = frcp(two_T);
    const float cross0 = fmaf(ox, dz, -__fmul_rn(z, dx));
    // centres c whose circle the line crosses: |cross0 + c*dx| < ra; and not wholly below the ray's start
    float c_lo = -LSCPM_INF, c_hi = LSCPM_INF;
    if (dx != 0.f) {
        const float u = __fmul_rn(__fsub_rn(-cross0, ra), inv_dx), v = __fmul_rn(__fadd_rn(-cross0, ra), inv_dx);
        c_lo = fminf(u, v); c_hi = fmaxf(u, v);
    } else if (!(fabsf(cross0) < ra)) {
        c_lo = LSCPM_INF; c_hi = -LSCPM_INF;
    }
    c_lo = fmaxf(c_lo, __fsub_rn(z, r_o));
    if (dz == 0.f) c_hi = fminf(c_hi, __fadd_rn(z, r_o));
    const float c_odd = __fadd_rn(T, delta);
    const float e1 = __fmul_rn(ceilf(__fmul_rn(c_lo, inv_2T)), two_T), e2 = __fadd_rn(e1, two_T);
    const float o1 = fmaf(ceilf(__fmul_rn(__fsub_rn(c_lo, c_odd), inv_2T)), two_T, c_odd), o2 = __fadd_rn(o1, two_T);
    const float first = fminf(e1, o1);
    const float second = fminf(fmaxf(e1, o1), fminf(e2, o2));
    ColumnHit r;
    r.t = image_hit(ox, z, dx, dz, cross0, ra2, r_o2, first, c_lo, c_hi);
    r.c = first;
    if (two) {
        const float t2 = image_hit(ox, z, dx, dz, cross0, ra2, r_o2, second, c_lo, c_hi);
        if (t2 < r.t) { r.t = t2; r.c = second; }
    }
    r.odd = (r.c == o1 || r.c == o2) ? 1 : 0;
    return r;
}

template <bool FOLD>
__device__ __forceinline__ void plate_flight(const DevScene& S, Photon& p, const float L, Hit& h, int& cls, int& bounces) {
    // mirror the problem so that the photon travels upwards: z -> s*z with s = sign(dz); the capillary axes then sit at s*zc
    const float sgn = p.dz >= 0.f ? 1.f : -1.f;
    const float z0 = __fmul_rn(sgn, p.z), dz = fabsf(p.dz), zc = __fmul_rn(sgn, S.zc);
    const float T = 2.f * S.hz;
    const float z_bot = __fsub_rn(-S.hz, zc);           // lower plate face relative to the capillary axes
    const float delta = -2.f * zc;                      // offset of the mirrored (odd) images of the axes
    // the real flat faces ahead
    const float ty = p.dy != 0.f ? __fmul_rn(__fsub_rn(copysignf(S.hy, p.dy), p.y), frcp(p.dy)) : LSCPM_INF;
    const float tzf = dz != 0.f ? __fmul_rn(__fsub_rn(__fsub_rn(S.hz, zc), z0), frcp(dz)) : LSCPM_INF;
    const float x_face = p.dx >= 0.f ? fmaf((float)(S.n_ch - 1 - p.kc), S.pitch, S.cell_hi_last)
                                     : fmaf(-(float)p.kc, S.pitch, S.cell_lo_first);
    const float tx = p.dx != 0.f ? __fmul_rn(__fsub_rn(x_face, p.xr), frcp(p.dx)) : LSCPM_INF;
    bool trapped = false;
    if (FOLD) {   // interface_eval's own test for the z faces
        const float n1 = S.med[MED_PLATE].n, n2 = S.med[MED_AIR].n;
        const float eta = __fmul_rn(n1, frcp(n2));
        const float c = fminf(dz, 1.f);
        const float s2 = fmaxf(fmaf(-c, c, 1.f), 0.f);
        trapped = n2 < n1 && fmaf(-__fmul_rn(eta, eta), s2, 1.f) < 0.f;
    }
    float t_end = ty; int face = p.dy > 0.f ? 3 : 2;
    if (tx < t_end) { t_end = tx; face = p.dx > 0.f ? 1 : 0; }
    if (!trapped && tzf < t_end) { t_end = tzf; face = p.dz > 0.f ? 5 : 4; }
    const bool absorbed = L < t_end;
    t_end = fminf(t_end, L);
    // capillary images: the photon's own column and the next one in the direction of travel.  A flight that reaches
    // beyond them stops at the far cell boundary (a virtual crossing, as between cells before: the next iteration draws
    // a new free path, which changes nothing because free paths are memoryless).  Two columns settle most flights and
    // keep the lanes of a warp in step; an open-ended column loop ran at 4 of 32 lanes.
    int hit_col = -1, hit_odd = 0; float hit_c = 0.f; float t_hit = t_end;
    bool crossing = false;
    const float a = fmaf(p.dx, p.dx, __fmul_rn(dz, dz));
    const int step = p.dx >= 0.f ? 1 : -1;
    if (S.has_caps && a > 1e-30f) {
        const float ra = __fmul_rn(S.r_o, fsqrt(a)), ra2 = __fmul_rn(a, S.r_o2);
        const float inv_dx = p.dx != 0.f ? frcp(p.dx) : 0.f;
        int j = p.kc;
#pragma unroll
        for (int trip = 0; trip < FLIGHT_COLUMNS; ++trip) {
            const float ox = fmaf(-(float)(j - p.kc), S.pitch, p.xr);
            if (__fsub_rn(fabsf(ox), S.r_o) > __fmul_rn(fabsf(p.dx), t_hit)) break;   // the column starts beyond the end of the flight
            const ColumnHit ch = column_images(ox, z0, p.dx, dz, ra, ra2, S.r_o, S.r_o2, inv_dx, T, delta, trip == 0);
            if (ch.t < t_hit) { t_hit = ch.t; hit_col = j; hit_c = ch.c; hit_odd = ch.odd; break; }
            // no hit in column j: is there another column ahead, and does the flight reach its cell?
            const int jn = j + step;
            if (p.dx == 0.f || jn < 0 || jn >= S.n_ch) break;
            if (trip == FLIGHT_COLUMNS - 1) {
                const float t_stop = __fmul_rn(__fsub_rn(__fmul_rn((float)(j - p.kc) + 0.5f * (float)step, S.pitch), p.xr), inv_dx);
                if (t_stop < t_hit) { t_hit = t_stop; crossing = true; hit_col = jn; }
                break;
            }
            j = jn;
        }
    }
    // move, then fold back into the real plate (selects, no branches)
    const float t = t_hit;
    const float x_rel = fmaf(p.dx, t, p.xr), zu = fmaf(dz, t, z0);
    p.y = fmaf(p.dy, t, p.y);
    const float inv_T = frcp(T);
    const bool image = hit_col >= 0 && !crossing;
    int k = p.kc + (int)rintf(__fmul_rn(x_rel, S.inv_pitch));
    k = min(max(k, 0), S.n_ch - 1);
    k = hit_col >= 0 ? hit_col : k;
    const float xr_new = fmaf(-(float)(k - p.kc), S.pitch, x_rel);
    p.xr = crossing ? (step > 0 ? -0.5f * S.pitch : 0.5f * S.pitch) : xr_new;   // a crossing lands on the boundary, in the cell ahead
    p.kc = k;
    const int m_img = (int)rintf(__fmul_rn(__fsub_rn(hit_c, hit_odd ? delta : 0.f), inv_T));
    const int m_flat = trapped ? (int)floorf(__fmul_rn(__fsub_rn(zu, z_bot), inv_T)) : 0;
    const int m = image ? m_img : m_flat;
    float z = image ? __fsub_rn(zu, hit_c) : fmaf(-(float)m, T, zu);
    float dz_out = dz;
    if (m & 1) { z = __fsub_rn(image ? 0.f : delta, z); dz_out = -dz; }
    p.z = __fmul_rn(sgn, z); p.dz = __fmul_rn(sgn, dz_out);
    bounces = m;                                        // the mirrored photon only ever moves up: m >= 0
    h.t = t; h.ch = p.kc; h.container = MED_PLATE;
    h.kind = image ? SURF_OUTER : (crossing ? SURF_CELL : SURF_BOX);
    h.face = image ? 0 : (crossing ? step : face);
    h.adj = image ? MED_OUTER : (crossing ? MED_PLATE : MED_AIR);
    cls = image ? CLS_CYL : (crossing ? CLS_CELL : (absorbed ? CLS_ABSORB : CLS_FLAT));
}

// first half of the iteration for a photon in the plate assembly (or waiting on its face): nearest interface, free
// path, move.  Returns -1 or BIN_KILL (step limit); `cls` names the interaction that follows, `h` describes it,
// `extra` counts the reflections folded into a plate flight.
// BRANCH_ENTRY: skip the geometry for a photon still waiting for its entry Fresnel test (pays when whole warps are in
// that state - the sorted kernel; the lane kernel computes and overrides, which costs no branch).
// FLIGHT: plate flights (the host enables them only when no outside box shares a z face of the plate: the fold's test
// is against the world's medium); FOLD: fold total internal reflections into them.
// LAZY (pool kernel): the photon record carries no attenuation coefficients; the container's is looked up from the
// wavelength here (same function, same value as refresh_alphas stored) and handed back in `alpha` for the absorption.
template <bool BRANCH_ENTRY, bool FLIGHT, bool FOLD, bool LAZY = false>
__device__ __forceinline__ int advance_hit(const DevScene& S, Photon& p, const uint32_t w_x, Hit& h, int& cls, int& extra, float& alpha_out) {
    extra = 0;
    cls = CLS_CELL;
    if (FLIGHT && p.medium == MED_PLATE) {
        const float alpha = LAZY ? medium_alpha(S, MED_PLATE, p.wl) : p.a_plate;
        alpha_out = alpha;
        const float depth = alpha <= ALPHA_ZERO ? LSCPM_INF : __fmul_rn(-flog(one_minus_u01(w_x)), frcp(alpha));
        int bounces;
        plate_flight<FOLD>(S, p, depth, h, cls, bounces);
        extra = bounces;
        p.steps += (cls != CLS_CELL ? 1 : 0) + bounces;
        return p.steps > MAXSTEPS ? BIN_KILL : -1;
    }
    if (BRANCH_ENTRY) {
        if (p.medium == MED_AIR) { h.t = 0.f; h.kind = SURF_BOX; h.face = p.aux & 7; h.container = MED_AIR; h.adj = MED_PLATE; }
        else h = compute_hit<!FLIGHT>(S, p);
    } else {
        h = compute_hit<!FLIGHT>(S, p);
        if (p.medium == MED_AIR) { h.t = 0.f; h.kind = SURF_BOX; h.face = p.aux & 7; h.container = MED_AIR; h.adj = MED_PLATE; }
    }
    // a virtual crossing is not a follow() iteration
    p.steps += h.kind != SURF_CELL ? 1 : 0;
    if (p.steps > MAXSTEPS) return BIN_KILL;
    // branch-free select of the container's attenuation (the compiler turned the ternary chain into a jump table)
    float alpha = 0.f;
    if (LAZY) {
        if (h.container == MED_PLATE || h.container == MED_OUTER || h.container == MED_INNER) alpha = medium_alpha(S, h.container, p.wl);
    } else {
        alpha = h.container == MED_PLATE ? p.a_plate : alpha;
        alpha = h.container == MED_OUTER ? p.a_outer : alpha;
        alpha = h.container == MED_INNER ? p.a_inner : alpha;
    }
    alpha_out = alpha;
    const float depth = alpha <= ALPHA_ZERO ? LSCPM_INF : fdiv(-flog(one_minus_u01(w_x)), alpha);
    const bool absorbed = depth < h.t;
    const float travel = absorbed ? depth : h.t;
    p.xr = fmaf(p.dx, travel, p.xr); p.y = fmaf(p.dy, travel, p.y); p.z = fmaf(p.dz, travel, p.z);
    cls = (h.kind == SURF_BOX || h.kind == SURF_CAP) ? CLS_FLAT : CLS_CYL;
    if (h.kind == SURF_CELL) cls = CLS_CELL;
    if (absorbed) cls = CLS_ABSORB;
    if (cls == CLS_CELL) {   // step into the neighbouring cell; nothing physical happens here
        p.kc += h.face;
        LSCPM_CHECK(p.kc >= 0 && p.kc < S.n_ch);
        p.xr = h.face > 0 ? -0.5f * S.pitch : 0.5f * S.pitch;
    }
    return -1;
}

template <bool BRANCH_ENTRY, bool FLIGHT, bool FOLD>
__device__ __forceinline__ int advance_hit(const DevScene& S, Photon& p, const uint32_t w_x, Hit& h, int& cls, int& extra) {
    float alpha;
    return advance_hit<BRANCH_ENTRY, FLIGHT, FOLD, false>(S, p, w_x, h, cls, extra, alpha);
}

// the sorted kernel's form: the result packed into the pending word that travels with the photon record
template <bool EXT, bool FLIGHT>
__device__ __forceinline__ int advance(const DevScene& S, Photon& p, const uint32_t w_x, unsigned& pend, unsigned& n_events) {
    if (EXT && p.medium == MED_OUT) { pend = CLS_OUT; return -1; }
    Hit h; int cls, extra;
    const int killed = advance_hit<true, FLIGHT, true>(S, p, w_x, h, cls, extra);
    if (EXT && FLIGHT) p.aux = note_folded(p.aux, extra);
    n_events += (unsigned)extra;
    pend = pack_pending(cls, h.kind, h.face, h.container, h.adj);
    return killed;
}

// absorbed inside the container: choose the component ~ alpha_i, radiative test.  Returns the tally bin, or -1 with
// p.aux carrying the emission table when the luminophore re-emits.
__device__ __forceinline__ int interact_absorb_alpha(const DevScene& S, Photon& p, int container, float alpha, float u_pick, uint32_t w_z, int& event);
__device__ __forceinline__ int interact_absorb(const DevScene& S, Photon& p, int container, float u_pick, uint32_t w_z, int& event) {
    float alpha = 0.f;
    alpha = container == MED_PLATE ? p.a_plate : alpha;
    alpha = container == MED_OUTER ? p.a_outer : alpha;
    alpha = container == MED_INNER ? p.a_inner : alpha;
    return interact_absorb_alpha(S, p, container, alpha, u_pick, w_z, event);
}
// the same with the container's attenuation coefficient handed in (pool kernel: advance_hit has just looked it up)
__device__ __forceinline__ int interact_absorb_alpha(const DevScene& S, Photon& p, int container, float alpha, float u_pick, uint32_t w_z, int& event) {
    const DevMedium& m = S.med[container];
    const int index = pick_component(S, container, p.wl, alpha, u_pick * alpha);
    const DevComponent& comp = m.comp[index];
    if (comp.kind == COMP_LUMINOPHORE && u01(w_z) < comp.qy) {
        p.aux = (p.aux & 0xffff) | (comp.ems_table << 16);
        event = EV_EMIT;
        return -1;
    }
    event = comp.kind == COMP_REACTOR ? EV_REACT : EV_ABSORB;
    LSCPM_CHECK(p.kc >= 0 && p.kc < S.n_ch && index >= 0 && index < m.n_comp);
    const int base = container == MED_PLATE ? S.abs_base_plate
                   : (container == MED_OUTER ? __ldg(S.abs_base_outer + p.kc) : __ldg(S.abs_base_inner + p.kc));
    return base + index;
}

// Fresnel test at the interface the photon has been moved to; returns the tally bin or -1
template <bool EXT>
__device__ __forceinline__ int interact_surface(const DevScene& S, Photon& p, int kind, int face, int container, int adj,
                                                float u_surf, int& event) {
    const bool flat = kind == SURF_BOX || kind == SURF_CAP;
    const int exit_offset = face >> 3;      // non-zero for a cap hit by the tube or the mixture (cap_tie)
    face &= 7;
    float nx, ny, nz;
    box_normal(face, nx, ny, nz);
    if (flat) {
        if (p.medium != MED_AIR) {   // snap the coordinate of the face that was hit
            if (face < 2) { if (kind == SURF_BOX) p.xr = (face & 1) ? S.cell_hi_last : S.cell_lo_first; }
            else if (face < 4) p.y = (face & 1) ? S.hy : -S.hy;
            else p.z = ((face & 1) ? S.hz : -S.hz) - S.zc;
        }
    } else {
        const float inv = frsqrt(fmaf(p.xr, p.xr, p.z * p.z));
        nx = p.xr * inv; ny = 0.f; nz = p.z * inv;
    }
    // leaving the plate the photon meets the world's medium even when an outside box lies straight ahead: adjacent is the
    // container of what lies beyond the face, and a box that is entered and left again is not one.  A box that SHARES the
    // face ties with it: when its entry sorts first it is the hit node and adjacent (tie_bits, evaluated on arrival)
    float n2 = S.med[adj].n;
    const bool leaving = flat && p.medium != MED_AIR;
    if (EXT && leaving && kind == SURF_BOX && S.ext_touch[face] != 0u && (tie_bits(p) & 2u)) {
        const int e = ext_touching(S, face, abs_x(S, p), p.y, p.z + S.zc, p.dx, p.dy, p.dz);
        if (e >= 0) n2 = S.ext[e].n;
    }
    const float n1 = S.med[container].n;
    const Interface I = interface_eval(p.dx, p.dy, p.dz, nx, ny, nz, n1, n2);
    const bool reflected = u_surf < I.R;
    if (reflected) reflect_dir(p, I); else refract_dir(p, I);
    event = reflected ? EV_REFLECT : EV_TRANSMIT;
    if (p.medium == MED_AIR) {
        if (reflected) return BIN_ENTRY_REFLECT;   // bounced off at first contact, leaves the scene
        p.medium = (p.aux >> 4) & 3;
        return -1;
    }
    if (EXT) {
        const bool cap_hit = kind == SURF_CAP && container != MED_PLATE;    // hit node (= container at a cap) is the tube or the mixture
        p.aux = note_surface(p.aux, false, cap_hit ? 7 : 0, face, leaving && !reflected);
        if (cap_hit && leaving && !reflected) p.aux = (p.aux & ~((0xff << 16) | (3 << 27))) | (min(p.kc, 255) << 16) | (container << 27);
    }
    if (reflected) return -1;
    if (flat) {
        if (EXT) {
            const float xa = abs_x(S, p);
            p.kc = 0; p.xr = xa - axis_x(S, 0);
            p.medium = MED_OUT;
            return -1;
        }
        p.medium = MED_AIR;
        return (p.steps + 1 > MAXSTEPS) ? BIN_KILL : S.exit_base_plate + face + exit_offset;
    }
    if (kind == SURF_OUTER) p.medium = (p.medium == MED_PLATE) ? MED_OUTER : MED_PLATE;
    else p.medium = (p.medium == MED_OUTER) ? MED_INNER : MED_OUTER;
    return -1;
}

// one whole iteration: advance, then the interaction it names
template <bool EXT, bool FLIGHT, bool FOLD = true>
__device__ __forceinline__ int photon_step(const DevScene& S, Photon& p, const uint4 w, StepOut& out) {
    out.extra = 0;
    if (EXT && p.medium == MED_OUT) {
        OutState q;
        q.x = abs_x(S, p); q.y = p.y; q.z = p.z + S.zc; q.dx = p.dx; q.dy = p.dy; q.dz = p.dz; q.aux = p.aux; q.steps = p.steps;
        const OutResult r = outside_step(S, q, u01(w.y));
        p.dx = r.dx; p.dy = r.dy; p.dz = r.dz; p.aux = r.aux; p.steps = r.steps;
        out.event = r.event; out.medium = MED_AIR; out.ch = r.node; out.kind = SURF_BOX; out.face = r.face; out.extra = r.extra;
        if (r.entered) { p.y = r.y; locate_in_plate(S, p, r.x, r.z); }
        else { p.kc = 0; p.xr = r.x - axis_x(S, 0); p.y = r.y; p.z = r.z - S.zc; }
        return r.bin;
    }
    out.ch = p.kc;
    Hit h; int cls;
    const int killed = advance_hit<false, FLIGHT, FOLD>(S, p, w.x, h, cls, out.extra);
    if (EXT && FLIGHT) p.aux = note_folded(p.aux, out.extra);
    out.medium = h.container; out.kind = h.kind; out.face = h.kind == SURF_CAP ? (h.face & 7) : h.face;
    if (killed >= 0) { out.event = EV_KILL; return killed; }
    if (cls == CLS_ABSORB) {
        const int bin = interact_absorb(S, p, h.container, u01(w.y), w.z, out.event);
        return bin < 0 ? STEP_EMIT : bin;
    }
    if (cls == CLS_CELL) { out.event = EV_NONE; return -1; }
    return interact_surface<EXT>(S, p, h.kind, h.face, h.container, h.adj, u01(w.y), out.event);
}

}  // namespace lscpm
