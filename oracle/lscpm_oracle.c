/*
 * CPU ORACLE (test infrastructure, NOT product code) — plain-C float64 restatement of the photon loop
 * the reference drives (miniplant/simulation_runner.py:38-66) through pvtrace 2.1
 * (fork cambiegroup/pvtrace@22c90761e2ee2a9b9193051f40ee553bd4f527f8, reference requirements.txt:43).
 *
 * PARITY PIN STATUS: per-ray values are parity unpinned against real pvtrace output (pvtrace is not vendored in
 * /root/reference and not installable offline); tallies are pinned statistically on the full-year results the
 * reference committed (tests/test_solar_stage.py).  This file is the fast twin of
 * oracle/pvtrace_oracle.py — the same generic algorithm (every geometry node intersected in its own
 * local frame, distance sort with exact ties in random order, container = nearest node hit exactly once, adjacent = the hit
 * node or, when the photon leaves its container, the container of what lies beyond (both pinned on reference output,
 * DESIGN.md section 5), Beer-Lambert free
 * path, Fresnel reflect/refract, kT-restricted re-emission), working on the flattened description of
 * include/lscpm.h instead of the scene graph.  tests/test_oracle_golden.py checks the two against
 * each other ray by ray; the Python twin is checked against the reference's own test bands and the
 * analytic known-answer vectors.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may load this.
 * Nothing here is specific to the LSC-PM layout: no plate-local frame, no capillary lookup — that is
 * the CUDA library's business, and this brute-force version is what it is checked against.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/lscpm.h"

#define ZERO_TOL 1e-12      /* "distance is zero / distances tie" in exact arithmetic (see pvtrace_oracle.py) */
#define KB_EV 8.617333262145e-5
#define HC_EV_NM 1240.8
#define MAXSTEPS 1000
#define MAX_HITS_PER_NODE 4

/* ------------------------------------------------------------------------------------------ */
/* RNG: xoshiro256** seeded per photon through splitmix64 — deliberately NOT the Philox stream of
 * the CUDA library, so the two implementations are statistically independent. */
typedef struct { uint64_t s[4]; } Rng;

static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rng_seed(Rng* r, uint64_t seed, uint64_t stream) {
    uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ULL + 0x2545F4914F6CDD1DULL);
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&x);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(Rng* r) {
    uint64_t* s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static double rng_uniform(Rng* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); } /* [0,1) */

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int geom, material;
    double size[3];
    double R[9], T[3];     /* local -> world rotation and translation */
} ONode;

typedef struct {
    const LscpmSceneDesc* d;
    int n_nodes;
    ONode* nodes;
    double** cdf;          /* per table: cumulative-sum CDF normalised by its maximum */
    int n_bins;
    int* exit_base;        /* per node */
    int* abs_base;         /* per node */
} OScene;

typedef struct { double t; int node, face; } Hit;

static int n_components_of(const LscpmSceneDesc* d, int node) {
    int m = d->node_material[node];
    return d->mat_comp_begin[m + 1] - d->mat_comp_begin[m];
}

static void oscene_free(OScene* s) {
    if (!s) return;
    if (s->cdf) { for (int t = 0; t < s->d->n_tables; ++t) free(s->cdf[t]); free(s->cdf); }
    free(s->nodes); free(s->exit_base); free(s->abs_base);
}

/* pvtrace Distribution: cdf = cumsum(y) / max(cdf) over the knots (spacing does not enter) */
static void build_cdf(const double* x, const double* y, int n, double* cdf) {
    (void)x;
    double acc = 0.0, peak = 0.0;
    for (int i = 0; i < n; ++i) { acc += y[i]; cdf[i] = acc; if (acc > peak) peak = acc; }
    if (peak > 0) for (int i = 0; i < n; ++i) cdf[i] /= peak;
}

static int oscene_init(OScene* s, const LscpmSceneDesc* d) {
    memset(s, 0, sizeof(*s));
    s->d = d;
    s->n_nodes = d->n_nodes;
    s->nodes = (ONode*)calloc((size_t)d->n_nodes, sizeof(ONode));
    s->exit_base = (int*)calloc((size_t)d->n_nodes, sizeof(int));
    s->abs_base = (int*)calloc((size_t)d->n_nodes, sizeof(int));
    s->cdf = (double**)calloc((size_t)(d->n_tables > 0 ? d->n_tables : 1), sizeof(double*));
    if (!s->nodes || !s->exit_base || !s->abs_base || !s->cdf) return LSCPM_ENOMEM;
    for (int i = 0; i < d->n_nodes; ++i) {
        ONode* n = &s->nodes[i];
        const double* p = d->node_pose + 16 * i;
        n->geom = d->node_geom[i];
        n->material = d->node_material[i];
        for (int k = 0; k < 3; ++k) n->size[k] = d->node_size[3 * i + k];
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) n->R[3 * r + c] = p[4 * r + c]; n->T[r] = p[4 * r + 3]; }
    }
    for (int t = 0; t < d->n_tables; ++t) {
        int lo = d->table_begin[t], n = d->table_begin[t + 1] - lo;
        s->cdf[t] = (double*)malloc(sizeof(double) * (size_t)n);
        if (!s->cdf[t]) return LSCPM_ENOMEM;
        build_cdf(d->table_x + lo, d->table_y + lo, n, s->cdf[t]);
    }
    /* bin layout: 0 kill, 1 exit/miss, 2 exit/entry_reflect, 6 face slots per non-world node,
     * then one slot per (node, component) — the contract of lscpm_b200/tally.py */
    int next = 3;
    for (int i = 1; i < d->n_nodes; ++i) { s->exit_base[i] = next; next += 6; }
    for (int i = 0; i < d->n_nodes; ++i) { s->abs_base[i] = next; next += n_components_of(d, i); }
    s->n_bins = next;
    return 0;
}

/* numpy.interp(v, xp, fp): clamped, segment = last j with xp[j] <= v */
static double interp(double v, const double* xp, const double* fp, int n) {
    if (v <= xp[0]) return fp[0];
    if (v >= xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1;              /* invariant: xp[lo] <= v < xp[hi] */
    while (hi - lo > 1) { int mid = (lo + hi) / 2; if (xp[mid] <= v) lo = mid; else hi = mid; }
    double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
    return slope * (v - xp[lo]) + fp[lo];
}

static double component_coefficient(const OScene* s, int comp, double wavelength) {
    const LscpmSceneDesc* d = s->d;
    int t = d->comp_abs_table[comp];
    if (t < 0) return d->comp_coefficient[comp];
    int lo = d->table_begin[t];
    return interp(wavelength, d->table_x + lo, d->table_y + lo, d->table_begin[t + 1] - lo);
}

/* ---- geometry in the node-local frame ------------------------------------------------------ */
static int box_hits(const double* size, const double* o, const double* dv, Hit* out) {
    double tmin = -INFINITY, tmax = INFINITY; int fmin = -1, fmax = -1;
    for (int ax = 0; ax < 3; ++ax) {
        double h = 0.5 * size[ax];
        if (dv[ax] == 0.0) { if (fabs(o[ax]) > h) return 0; continue; }
        double tlo = (-h - o[ax]) / dv[ax], thi = (h - o[ax]) / dv[ax];
        int flo = 2 * ax, fhi = 2 * ax + 1;
        if (tlo > thi) { double tt = tlo; tlo = thi; thi = tt; int ff = flo; flo = fhi; fhi = ff; }
        if (tlo > tmin) { tmin = tlo; fmin = flo; }
        if (thi < tmax) { tmax = thi; fmax = fhi; }
    }
    if (tmin > tmax) return 0;
    int n = 0;
    if (tmin >= 0.0) { out[n].t = tmin; out[n].face = fmin; ++n; }
    if (tmax >= 0.0) { out[n].t = tmax; out[n].face = fmax; ++n; }
    return n;
}

static int quadratic(double a, double b, double c, double* r) {
    if (a == 0.0) return 0;
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return 0;
    double root = sqrt(disc);
    double q = -0.5 * (b + copysign(root, b));
    if (q == 0.0) { r[0] = r[1] = 0.0; return 2; }
    double t1 = q / a, t2 = c / q;
    if (t1 > t2) { double tt = t1; t1 = t2; t2 = tt; }
    r[0] = t1; r[1] = t2;
    return 2;
}

static int cylinder_hits(const double* size, const double* o, const double* dv, Hit* out) {
    double half = 0.5 * size[0], radius = size[1], r[2];
    int n = 0;
    int nr = quadratic(dv[0] * dv[0] + dv[1] * dv[1], 2.0 * (dv[0] * o[0] + dv[1] * o[1]),
                       o[0] * o[0] + o[1] * o[1] - radius * radius, r);
    for (int i = 0; i < nr; ++i)
        if (r[i] >= 0.0 && fabs(o[2] + r[i] * dv[2]) <= half) { out[n].t = r[i]; out[n].face = 0; ++n; }
    if (dv[2] != 0.0) {
        for (int face = 1; face <= 2; ++face) {
            double zc = face == 1 ? -half : half, t = (zc - o[2]) / dv[2];
            if (t >= 0.0) {
                double x = o[0] + t * dv[0], y = o[1] + t * dv[1];
                if (x * x + y * y < radius * radius) { out[n].t = t; out[n].face = face; ++n; }
            }
        }
    }
    return n;
}

static int sphere_hits(const double* size, const double* o, const double* dv, Hit* out) {
    double r[2];
    int nr = quadratic(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2], 2.0 * (dv[0] * o[0] + dv[1] * o[1] + dv[2] * o[2]),
                       o[0] * o[0] + o[1] * o[1] + o[2] * o[2] - size[0] * size[0], r);
    int n = 0;
    for (int i = 0; i < nr; ++i) if (r[i] >= 0.0) { out[n].t = r[i]; out[n].face = 0; ++n; }
    return n;
}

static void to_local(const ONode* n, const double* p, const double* v, double* o, double* dv) {
    double q[3] = {p[0] - n->T[0], p[1] - n->T[1], p[2] - n->T[2]};
    for (int c = 0; c < 3; ++c) {      /* R^T */
        o[c] = n->R[c] * q[0] + n->R[3 + c] * q[1] + n->R[6 + c] * q[2];
        dv[c] = n->R[c] * v[0] + n->R[3 + c] * v[1] + n->R[6 + c] * v[2];
    }
}

/* All intersections sorted by distance.  Exact ties (coincident faces, within ZERO_TOL): pvtrace sorts distances that each
 * node computed in its own frame, so float64 round-off decides their order - restated as a uniformly random order when a
 * generator is given (the photon loop), node (pre-order) order otherwise (probes). */
static int scene_intersections(const OScene* s, const double* p, const double* v, Hit* out, Rng* rng) {
    int total = 0;
    Hit local[MAX_HITS_PER_NODE];
    for (int i = 0; i < s->n_nodes; ++i) {
        const ONode* n = &s->nodes[i];
        double o[3], dv[3];
        to_local(n, p, v, o, dv);
        int k = n->geom == LSCPM_GEOM_BOX ? box_hits(n->size, o, dv, local)
              : n->geom == LSCPM_GEOM_CYLINDER ? cylinder_hits(n->size, o, dv, local) : sphere_hits(n->size, o, dv, local);
        for (int j = 0; j < k; ++j) {
            Hit h = local[j]; h.node = i;
            int pos = total;
            while (pos > 0) {
                const Hit* prev = &out[pos - 1];
                if (prev->t > h.t + ZERO_TOL || (fabs(prev->t - h.t) <= ZERO_TOL && prev->node > h.node)) { out[pos] = *prev; --pos; }
                else break;
            }
            out[pos] = h; ++total;
        }
    }
    if (rng) {
        int start = 0;
        for (int end = 1; end <= total; ++end) {
            if (end == total || out[end].t - out[start].t > ZERO_TOL) {
                for (int i = end - 1; i > start; --i) {           /* Fisher-Yates over the tied group */
                    int j = start + (int)(rng_uniform(rng) * (double)(i - start + 1));
                    if (j > i) j = i;
                    Hit tmp = out[i]; out[i] = out[j]; out[j] = tmp;
                }
                start = end;
            }
        }
    }
    return total;
}

typedef struct { int hit, container, adjacent, face; double distance; } NextHit;

/* pvtrace find_container: the nearest of the nodes that appear exactly once in hits[0..n) (the ray is inside them) */
static int find_container(const OScene* s, const Hit* hits, int n, int* counts) {
    if (n == 1) return hits[0].node;
    memset(counts, 0, sizeof(int) * (size_t)s->n_nodes);
    for (int i = 0; i < n; ++i) counts[hits[i].node]++;
    for (int i = 0; i < n; ++i) if (counts[hits[i].node] == 1) return hits[i].node;
    return -1;
}

/* literal next_hit; returns 0 when the ray leaves the scene */
static int next_hit(const OScene* s, const double* p, const double* v, Hit* scratch, int* counts, NextHit* nh, Rng* rng) {
    int n_all = scene_intersections(s, p, v, scratch, rng), n = 0;
    for (int i = 0; i < n_all; ++i) if (!(fabs(scratch[i].t) <= ZERO_TOL)) scratch[n++] = scratch[i];
    if (n == 0) return 0;
    nh->hit = scratch[0].node; nh->face = scratch[0].face; nh->distance = scratch[0].t;
    if (n == 1) { nh->container = nh->hit; nh->adjacent = -1; return 1; }
    nh->container = find_container(s, scratch, n, counts);
    if (nh->container < 0) nh->container = nh->hit;   /* cannot happen inside the world sphere */
    /* leaving the container: adjacent = container of what lies beyond the surface (the rule applied to the rest) */
    nh->adjacent = nh->container == nh->hit ? find_container(s, scratch + 1, n - 1, counts) : nh->hit;
    if (nh->adjacent < 0) nh->adjacent = 0;
    return 1;
}

static void surface_normal_world(const OScene* s, int node, int face, const double* point, double* nw) {
    const ONode* n = &s->nodes[node];
    double o[3], unused[3], zero[3] = {0, 0, 0}, nl[3] = {0, 0, 0};
    to_local(n, point, zero, o, unused);
    if (n->geom == LSCPM_GEOM_BOX) nl[face / 2] = (face % 2) ? 1.0 : -1.0;
    else if (n->geom == LSCPM_GEOM_CYLINDER) {
        if (face == 0) { double r = hypot(o[0], o[1]); nl[0] = o[0] / r; nl[1] = o[1] / r; }
        else nl[2] = face == 1 ? -1.0 : 1.0;
    } else { double r = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]); for (int k = 0; k < 3; ++k) nl[k] = o[k] / r; }
    for (int r = 0; r < 3; ++r) nw[r] = n->R[3 * r] * nl[0] + n->R[3 * r + 1] * nl[1] + n->R[3 * r + 2] * nl[2];
}

/* ---- optics --------------------------------------------------------------------------------- */
static double fresnel_reflectivity(double angle, double n1, double n2) {
    if (n2 < n1 && angle > asin(n2 / n1)) return 1.0;
    double c = cos(angle), sn = sin(angle);
    double k2 = 1.0 - (n1 / n2 * sn) * (n1 / n2 * sn);
    double k = sqrt(k2 > 0 ? k2 : 0.0);
    double rs = (n1 * c - n2 * k) / (n1 * c + n2 * k), rp = (n1 * k - n2 * c) / (n1 * k + n2 * c);
    return 0.5 * (rs * rs + rp * rp);
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

typedef struct { double normal[3], reflectivity, reflected[3], refracted[3]; int tir; } Interface;

static void interface_eval(const double* dir, const double* raw_normal, double n1, double n2, Interface* out) {
    double nrm[3] = {raw_normal[0], raw_normal[1], raw_normal[2]};
    if (dot3(nrm, dir) < 0.0) for (int k = 0; k < 3; ++k) nrm[k] = -nrm[k];
    double c = dot3(nrm, dir);
    if (c > 1.0) c = 1.0; if (c < -1.0) c = -1.0;
    out->reflectivity = fresnel_reflectivity(acos(c), n1, n2);
    out->tir = out->reflectivity >= 1.0;
    double dn = dot3(dir, nrm), n = n1 / n2;
    double cc2 = 1.0 - n * n * (1.0 - dn * dn), cc = sqrt(cc2 > 0 ? cc2 : 0.0);
    double sign = dn < 0.0 ? -1.0 : 1.0;
    for (int k = 0; k < 3; ++k) {
        out->normal[k] = nrm[k];
        out->reflected[k] = dir[k] - 2.0 * dn * nrm[k];
        out->refracted[k] = out->tir ? 0.0 : n * dir[k] + sign * (cc - sign * n * dn) * nrm[k];
    }
}

/* ---- photon loop ---------------------------------------------------------------------------- */
typedef struct {
    int max_steps; int32_t* n_steps; int32_t* ev; int32_t* node; double* pos; double* dir; double* wl;
} Recorder;

static void record(const Recorder* rec, int64_t photon, int* k, int event, int node, const double* p, const double* v, double wl) {
    if (!rec || *k >= rec->max_steps) { if (rec) ++*k; return; }
    int64_t at = photon * rec->max_steps + *k;
    rec->ev[at] = event; rec->node[at] = node; rec->wl[at] = wl;
    for (int c = 0; c < 3; ++c) { rec->pos[3 * at + c] = p[c]; rec->dir[3 * at + c] = v[c]; }
    ++*k;
}

/* Follows one photon; returns its tally bin and adds to the event counters. */
static int follow(const OScene* s, double* p, double* v, double wl, Rng* rng, Hit* scratch, int* counts,
                  int64_t* n_events, int64_t* n_emit, const Recorder* rec, int64_t photon) {
    const LscpmSceneDesc* d = s->d;
    int count = 0, k = 0, n_surface = 0, emitted = 0;
    int last_node = -1, last_face = -1, last_event = -1, last_container = -1;
    double coefs[16];
    record(rec, photon, &k, LSCPM_EV_GENERATE, -1, p, v, wl);
    int bin = -1;
    for (;;) {
        ++count;
        if (count > MAXSTEPS) { record(rec, photon, &k, LSCPM_EV_KILL, -1, p, v, wl); bin = 0; break; }
        NextHit nh;
        int found = next_hit(s, p, v, scratch, counts, &nh, rng);
        if (found && nh.hit == 0) for (int c = 0; c < 3; ++c) p[c] += nh.distance * v[c];
        if (!found || nh.hit == 0) {
            record(rec, photon, &k, LSCPM_EV_EXIT, found ? 0 : -1, p, v, wl);
            if (n_surface == 0) bin = 1;
            else if (n_surface == 1 && last_event == LSCPM_EV_REFLECT && last_container == 0 && !emitted) bin = 2;
            else bin = s->exit_base[last_node] + last_face;
            break;
        }
        int mat = d->node_material[nh.container];
        int c0 = d->mat_comp_begin[mat], nc = d->mat_comp_begin[mat + 1] - c0;
        double alpha = 0.0;
        for (int c = 0; c < nc; ++c) { coefs[c] = component_coefficient(s, c0 + c, wl); alpha += coefs[c]; }
        double depth;
        if (fabs(alpha) <= 1e-8) depth = INFINITY;            /* numpy.isclose(alpha, 0) */
        else if (!isfinite(alpha)) depth = 0.0;
        else depth = -log(1.0 - rng_uniform(rng)) / alpha;
        if (depth < nh.distance) {
            for (int c = 0; c < 3; ++c) p[c] += depth * v[c];
            double u = rng_uniform(rng) * alpha, acc = 0.0;
            int index = nc - 1;
            for (int c = 0; c < nc; ++c) { acc += coefs[c]; if (u < acc) { index = c; break; } }
            int comp = c0 + index, kind = d->comp_kind[comp];
            ++*n_events;
            if (kind == LSCPM_COMP_LUMINOPHORE && rng_uniform(rng) < d->comp_quantum_yield[comp]) {
                double phi = 2.0 * M_PI * rng_uniform(rng), mu = 2.0 * rng_uniform(rng) - 1.0, theta = acos(mu);
                v[0] = sin(theta) * cos(phi); v[1] = sin(theta) * sin(phi); v[2] = cos(theta);
                int t = d->comp_ems_table[comp], lo = d->table_begin[t], n = d->table_begin[t + 1] - lo;
                double ev = HC_EV_NM / wl + 1.5 * KB_EV * 300.0;
                double p1 = interp(HC_EV_NM / ev, d->table_x + lo, s->cdf[t], n);
                double gamma = p1 + (1.0 - p1) * rng_uniform(rng);
                wl = interp(gamma, s->cdf[t], d->table_x + lo, n);
                ++*n_emit; emitted = 1;
                record(rec, photon, &k, LSCPM_EV_EMIT, nh.container, p, v, wl);
                continue;
            }
            record(rec, photon, &k, kind == LSCPM_COMP_REACTOR ? LSCPM_EV_REACT : LSCPM_EV_ABSORB, nh.container, p, v, wl);
            bin = s->abs_base[nh.container] + index;
            break;
        }
        for (int c = 0; c < 3; ++c) p[c] += nh.distance * v[c];
        double raw[3];
        surface_normal_world(s, nh.hit, nh.face, p, raw);
        Interface itf;
        interface_eval(v, raw, d->mat_refractive_index[d->node_material[nh.container]],
                       d->mat_refractive_index[d->node_material[nh.adjacent]], &itf);
        int reflected = rng_uniform(rng) < itf.reflectivity;
        for (int c = 0; c < 3; ++c) v[c] = reflected ? itf.reflected[c] : itf.refracted[c];
        ++*n_events; ++n_surface;
        last_node = nh.hit; last_face = nh.face; last_container = nh.container;
        last_event = reflected ? LSCPM_EV_REFLECT : LSCPM_EV_TRANSMIT;
        record(rec, photon, &k, last_event, nh.hit, p, v, wl);
    }
    if (rec) rec->n_steps[photon] = k;
    return bin;
}

/* ---- light (reference miniplant/utils.py:87-142, scene_creator.py:288-304) -------------------- */
typedef struct { const double* x; const double* cdf; int n; } Spectrum;   /* cdf: cumulative sum, see build_cdf */

static void emit_ray(const LscpmBatchDesc* b, const Spectrum* sp, Rng* rng, double* p, double* v, double* wl) {
    double mx = (2.0 * rng_uniform(rng) - 1.0) * b->mask_half[0], my = (2.0 * rng_uniform(rng) - 1.0) * b->mask_half[1];
    double anchor[3];
    for (int r = 0; r < 3; ++r) anchor[r] = b->mask_pose[4 * r] * mx + b->mask_pose[4 * r + 1] * my + b->mask_pose[4 * r + 3];
    if (b->light_kind == LSCPM_LIGHT_DIRECT) {
        for (int c = 0; c < 3; ++c) v[c] = b->direction[c];
    } else {
        double tilt = b->tilt_deg * M_PI / 180.0, zen, az;
        do {
            az = rng_uniform(rng) * 360.0 * M_PI / 180.0;
            zen = rng_uniform(rng) * 90.0 * M_PI / 180.0;
        } while (cos(tilt) * cos(zen) + sin(tilt) * sin(zen) * cos(az) < 0.0);
        v[0] = -sin(zen) * cos(az); v[1] = -sin(zen) * sin(az); v[2] = -cos(zen);
    }
    for (int c = 0; c < 3; ++c) p[c] = anchor[c] - b->back_off * v[c];
    *wl = sp->n > 0 ? interp(rng_uniform(rng), sp->cdf, sp->x, sp->n) : b->wavelength_nm;
}

static int spectrum_prepare(const LscpmBatchDesc* b, const LscpmSpectraDesc* spectra, Spectrum* sp, double** owned) {
    sp->n = 0; sp->x = NULL; sp->cdf = NULL; *owned = NULL;
    if (b->spectrum < 0) return 0;
    if (!spectra || b->spectrum >= spectra->n_spectra) return LSCPM_EINVAL;
    int lo = spectra->begin[b->spectrum], n = spectra->begin[b->spectrum + 1] - lo;
    double* cdf = (double*)malloc(sizeof(double) * (size_t)n);
    if (!cdf) return LSCPM_ENOMEM;
    build_cdf(spectra->x + lo, spectra->y + lo, n, cdf);
    sp->x = spectra->x + lo; sp->cdf = cdf; sp->n = n; *owned = cdf;
    return 0;
}

/* ---- exported entry points (loaded with ctypes by tests / bench) ------------------------------ */
int oracle_n_bins(const LscpmSceneDesc* d) {
    OScene s; int rc = oscene_init(&s, d); int n = rc ? rc : s.n_bins; oscene_free(&s); return n;
}

/* Traces photons [photon_begin, photon_begin+photon_count) of ONE batch; tallies[n_bins+2] is ADDED to. */
int oracle_trace(const LscpmSceneDesc* d, const LscpmBatchDesc* batch, const LscpmSpectraDesc* spectra,
                 uint64_t photon_begin, uint64_t photon_count, uint64_t seed, int n_threads, int64_t* tallies) {
    OScene s; int rc = oscene_init(&s, d);
    if (rc) { oscene_free(&s); return rc; }
    Spectrum sp; double* owned;
    rc = spectrum_prepare(batch, spectra, &sp, &owned);
    if (rc) { oscene_free(&s); return rc; }
    int nb = s.n_bins + LSCPM_N_COUNTERS;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    int failed = 0;
#pragma omp parallel
    {
        int64_t* local = (int64_t*)calloc((size_t)nb, sizeof(int64_t));
        Hit* scratch = (Hit*)malloc(sizeof(Hit) * (size_t)(s.n_nodes * MAX_HITS_PER_NODE));
        int* counts = (int*)malloc(sizeof(int) * (size_t)s.n_nodes);
        if (!local || !scratch || !counts) {
#pragma omp atomic write
            failed = 1;
        } else {
#pragma omp for schedule(dynamic, 256)
            for (int64_t i = 0; i < (int64_t)photon_count; ++i) {
                Rng rng; rng_seed(&rng, seed ^ ((uint64_t)batch->batch_id << 40), photon_begin + (uint64_t)i);
                double p[3], v[3], wl;
                emit_ray(batch, &sp, &rng, p, v, &wl);
                int bin = follow(&s, p, v, wl, &rng, scratch, counts, &local[s.n_bins], &local[s.n_bins + 1], NULL, 0);
                local[bin]++;
            }
#pragma omp critical
            for (int b = 0; b < nb; ++b) tallies[b] += local[b];
        }
        free(local); free(scratch); free(counts);
    }
    free(owned); oscene_free(&s);
    return failed ? LSCPM_ENOMEM : 0;
}

int oracle_emit(const LscpmBatchDesc* batch, const LscpmSpectraDesc* spectra, uint64_t photon_begin, int64_t n,
                uint64_t seed, double* pos, double* dir, double* wavelength) {
    Spectrum sp; double* owned;
    int rc = spectrum_prepare(batch, spectra, &sp, &owned);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        Rng rng; rng_seed(&rng, seed ^ ((uint64_t)batch->batch_id << 40), photon_begin + (uint64_t)i);
        emit_ray(batch, &sp, &rng, pos + 3 * i, dir + 3 * i, wavelength + i);
    }
    free(owned);
    return 0;
}

int oracle_probe(const LscpmSceneDesc* d, int64_t n, const double* pos, const double* dir,
                 int32_t* hit, int32_t* container, int32_t* adjacent, int32_t* face,
                 double* distance, double* point, double* normal, double* reflectivity,
                 double* reflected, double* refracted) {
    OScene s; int rc = oscene_init(&s, d);
    if (rc) { oscene_free(&s); return rc; }
    Hit* scratch = (Hit*)malloc(sizeof(Hit) * (size_t)(s.n_nodes * MAX_HITS_PER_NODE));
    int* counts = (int*)malloc(sizeof(int) * (size_t)s.n_nodes);
    for (int64_t i = 0; i < n; ++i) {
        NextHit nh;
        const double* p = pos + 3 * i; const double* v = dir + 3 * i;
        hit[i] = container[i] = adjacent[i] = face[i] = -1;
        distance[i] = reflectivity[i] = 0.0;
        for (int c = 0; c < 3; ++c) point[3 * i + c] = normal[3 * i + c] = reflected[3 * i + c] = refracted[3 * i + c] = 0.0;
        if (!next_hit(&s, p, v, scratch, counts, &nh, NULL)) continue;
        hit[i] = nh.hit; container[i] = nh.container; adjacent[i] = nh.adjacent; face[i] = nh.face; distance[i] = nh.distance;
        for (int c = 0; c < 3; ++c) point[3 * i + c] = p[c] + nh.distance * v[c];
        if (nh.adjacent < 0 || nh.hit == 0) continue;
        double raw[3]; Interface itf;
        surface_normal_world(&s, nh.hit, nh.face, point + 3 * i, raw);
        interface_eval(v, raw, d->mat_refractive_index[d->node_material[nh.container]],
                       d->mat_refractive_index[d->node_material[nh.adjacent]], &itf);
        reflectivity[i] = itf.reflectivity;
        for (int c = 0; c < 3; ++c) { normal[3 * i + c] = itf.normal[c]; reflected[3 * i + c] = itf.reflected[c]; refracted[3 * i + c] = itf.refracted[c]; }
    }
    free(scratch); free(counts); oscene_free(&s);
    return 0;
}

int oracle_follow(const LscpmSceneDesc* d, int64_t n, const double* pos, const double* dir, const double* wavelength,
                  uint32_t batch_id, uint64_t photon_begin, uint64_t seed, int32_t max_steps,
                  int32_t* n_steps, int32_t* step_event, int32_t* step_node, double* step_pos, double* step_dir,
                  double* step_wavelength, int32_t* final_bin) {
    OScene s; int rc = oscene_init(&s, d);
    if (rc) { oscene_free(&s); return rc; }
    Hit* scratch = (Hit*)malloc(sizeof(Hit) * (size_t)(s.n_nodes * MAX_HITS_PER_NODE));
    int* counts = (int*)malloc(sizeof(int) * (size_t)s.n_nodes);
    Recorder rec = {max_steps, n_steps, step_event, step_node, step_pos, step_dir, step_wavelength};
    int64_t ev = 0, em = 0;
    for (int64_t i = 0; i < n; ++i) {
        Rng rng; rng_seed(&rng, seed ^ ((uint64_t)batch_id << 40), photon_begin + (uint64_t)i);
        double p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, v[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
        final_bin[i] = follow(&s, p, v, wavelength[i], &rng, scratch, counts, &ev, &em, max_steps > 0 ? &rec : NULL, i);
    }
    free(scratch); free(counts); oscene_free(&s);
    return 0;
}
