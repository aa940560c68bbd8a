/*
 * lscpm.h — C-ABI of the B200-native LSC-PM photon tracer (liblscpm.so).
 *
 * Drop-in boundary for the hot path of dcambie/LSC-PM_solar_miniplant.  The reference is pure Python;
 * its seam to the ray-tracing engine is the pvtrace object API consumed in
 * miniplant/simulation_runner.py:38-66.  Each entry point below cites the reference interface it
 * replaces.  Plain pointers and sizes only: no torch / Python types cross this boundary.
 *
 * Ownership: the caller owns every buffer it passes in; the library owns only the opaque
 * LscpmScene handle (device copies of the lowered scene) until lscpm_scene_destroy.
 * Errors: every function returns 0 on success or a negative LSCPM_E* code; the message of the last
 * failure on the calling thread is available from lscpm_last_error().  Nothing throws or exits.
 * Threading: a scene handle is immutable after creation; concurrent launches on different streams
 * are allowed.  One host thread (or process / rank) per GPU.
 */
#ifndef LSCPM_H
#define LSCPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSCPM_ABI_VERSION 1

/* error codes */
#define LSCPM_OK 0
#define LSCPM_EINVAL (-1)      /* malformed description / argument */
#define LSCPM_EUNSUPPORTED (-2)/* scene is valid pvtrace but outside the LSC-PM family the kernels handle */
#define LSCPM_ECUDA (-3)       /* CUDA runtime error (no device, launch failure, ...) */
#define LSCPM_ENOMEM (-4)

/* geometry kinds (pvtrace Sphere / Box / Cylinder; reference miniplant/scene_creator.py:58-61,87-97,219-246) */
#define LSCPM_GEOM_SPHERE 0   /* size = (radius, -, -)                       */
#define LSCPM_GEOM_BOX 1      /* size = full edge lengths (sx, sy, sz)        */
#define LSCPM_GEOM_CYLINDER 2 /* size = (length, radius, -), axis = local z   */

/* material component kinds (pvtrace Absorber / Luminophore and the fork's Reactor;
 * reference miniplant/scene_creator.py:67,72-84,216) */
#define LSCPM_COMP_ABSORBER 0
#define LSCPM_COMP_LUMINOPHORE 1
#define LSCPM_COMP_REACTOR 2

/* light kinds: reference miniplant/scene_creator.py:277-311 (direct) and :314-331 + utils.py:102-142 (diffuse) */
#define LSCPM_LIGHT_DIRECT 0
#define LSCPM_LIGHT_DIFFUSE 1

/* terminal / step events, numbered like oracle/ (pvtrace Event plus the fork's REACT) */
#define LSCPM_EV_GENERATE 0
#define LSCPM_EV_REFLECT 1
#define LSCPM_EV_TRANSMIT 2
#define LSCPM_EV_ABSORB 3
#define LSCPM_EV_EMIT 4
#define LSCPM_EV_EXIT 5
#define LSCPM_EV_KILL 6
#define LSCPM_EV_REACT 7

/*
 * Flattened scene graph (what scene_creator.flatten() emits).  Geometry nodes only, in scene-graph
 * pre-order (parents before children, children in insertion order).  Node 0 is the world node.
 * Coincident faces (capillary end caps in the plate's faces, boxes sharing a plate face): the photon
 * loop orders such tied intersections at random, as float64 round-off does in pvtrace; lscpm_probe
 * reports them in this pre-order.  Replaces the pvtrace Scene object handed to
 * _common_simulation_runner (reference miniplant/simulation_runner.py:16-21).
 */
typedef struct LscpmSceneDesc {
    int32_t n_nodes;
    const int32_t* node_geom;        /* [n_nodes] LSCPM_GEOM_*                                          */
    const int32_t* node_parent;      /* [n_nodes] nearest geometry ancestor, -1 for node 0              */
    const int32_t* node_material;    /* [n_nodes] index into the material arrays                        */
    const double* node_size;         /* [n_nodes*3]                                                     */
    const double* node_pose;         /* [n_nodes*16] row-major 4x4, node-local -> world (rigid)         */
    const char* const* node_name;    /* [n_nodes] NUL-terminated UTF-8, used for tally bin names        */

    int32_t n_materials;
    const double* mat_refractive_index; /* [n_materials]                                                */
    const int32_t* mat_comp_begin;      /* [n_materials+1] CSR offsets into the component arrays        */

    int32_t n_components;
    const int32_t* comp_kind;           /* [n_components] LSCPM_COMP_*                                  */
    const double* comp_coefficient;     /* [n_components] constant absorption coefficient, m^-1         */
    const int32_t* comp_abs_table;      /* [n_components] table id of a tabulated coefficient, or -1    */
    const int32_t* comp_ems_table;      /* [n_components] table id of the emission spectrum, or -1      */
    const double* comp_quantum_yield;   /* [n_components]                                               */

    int32_t n_tables;
    const int32_t* table_begin;         /* [n_tables+1] CSR offsets into the knot arrays                */
    const double* table_x;              /* knots: wavelength in nm, increasing                          */
    const double* table_y;              /* knots: value (m^-1 or arbitrary units)                       */
} LscpmSceneDesc;

/*
 * One batch = one light configuration (one time point of the drivers: reference
 * miniplant/full_simulation.py:64-102, angle_optimization_simulation.py:53-84).
 * Anchor points: uniform on the rectangle |x|<=mask_half[0], |y|<=mask_half[1], z=0 of the mask frame,
 * mapped to the world by mask_pose (reference utils.py:87-99).  A ray starts at anchor - back_off*d and
 * travels along the unit vector d, so it passes through its anchor.
 * DIRECT: d = direction (the negated sun vector; the light node sits back_off = 1 m up-sun,
 * scene_creator.py:288-304).  DIFFUSE: d = -sky with sky from zenith~U(0,90) deg, azimuth~U(0,360) deg,
 * redrawn while cos(t)cos(z)+sin(t)sin(z)cos(az) < 0, t = tilt_deg (utils.py:102-122,135-142).
 * Wavelength: constant wavelength_nm when spectrum < 0, else inverse-CDF sample of launch spectrum
 * table `spectrum` (pvtrace Distribution: cdf = cumsum(y)/max over the knots, reference utils.py:39-46).
 */
typedef struct LscpmBatchDesc {
    int32_t light_kind;
    int32_t spectrum;
    uint32_t batch_id;       /* global batch number: part of the RNG counter, so results do not depend on sharding */
    uint32_t reserved;
    double wavelength_nm;
    double mask_half[2];
    double mask_pose[12];    /* row-major 3x4: mask frame -> world */
    double direction[3];     /* DIRECT: world unit vector the photons travel along */
    double back_off;         /* metres behind the anchor where the ray starts */
    double tilt_deg;         /* DIFFUSE: tilt of the front-face rejection test */
} LscpmBatchDesc;

/* Launch-level spectrum tables (the per-time-point 27-knot photon-flux distributions of
 * reference miniplant/solar_data.py:94-111), CSR like the scene tables. */
typedef struct LscpmSpectraDesc {
    int32_t n_spectra;
    const int32_t* begin;    /* [n_spectra+1] */
    const double* x;         /* wavelength knots, nm */
    const double* y;         /* non-negative weights  */
} LscpmSpectraDesc;

typedef struct LscpmScene LscpmScene;

/* ---- library ------------------------------------------------------------------------------- */
int lscpm_abi_version(void);
const char* lscpm_last_error(void);
/* The counter-based generator the kernels use, evaluated on the host (Philox4x32-10; Random123 known-answer
 * vectors apply): out = philox(counter, key).  Pure function, no device needed. */
void lscpm_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);
/* number of visible CUDA devices (0 without a GPU); never fails */
int lscpm_device_count(void);

/* ---- scene handle -------------------------------------------------------------------------- */
/* Lower a flattened scene for `device` and upload it.  Replaces building/pickling the pvtrace Scene
 * per call (reference miniplant/simulation_runner.py:97-105).  LSCPM_EUNSUPPORTED when the graph is
 * not world > plate box > parallel capillary cylinders [> coaxial inner cylinder]. */
int lscpm_scene_create(const LscpmSceneDesc* desc, int device, LscpmScene** out);
int lscpm_scene_destroy(LscpmScene* scene);

/* Tally layout.  n_bins counts photon bins (their sum equals the photons traced) and excludes the
 * LSCPM_N_COUNTERS trailing event counters stored after them in every tally row. */
#define LSCPM_N_COUNTERS 2      /* [n_bins+0] interaction events, [n_bins+1] emission events */
int lscpm_scene_n_bins(const LscpmScene* scene);
/* Writes the NUL-terminated name of bin `bin` into buf (truncating); returns its full length. */
int lscpm_scene_bin_name(const LscpmScene* scene, int bin, char* buf, int buf_len);
/* Which trace kernel launches on this scene use (chosen at lscpm_scene_create from the scene family and shared-memory
 * needs), as text for benchmark records: writes the NUL-terminated name, returns its length. */
int lscpm_scene_kernel_name(const LscpmScene* scene, char* buf, int buf_len);
/* Same layout computed from a description alone (no device needed). */
int lscpm_desc_n_bins(const LscpmSceneDesc* desc);
int lscpm_desc_bin_name(const LscpmSceneDesc* desc, int bin, char* buf, int buf_len);

/* ---- the hot path -------------------------------------------------------------------------- */
/*
 * Trace photons [photon_begin, photon_begin+photon_count) of every batch to termination and ADD their
 * terminal classification into tallies_dev[b*(n_bins+LSCPM_N_COUNTERS) + bin] (int64, device memory,
 * caller-zeroed).  Asynchronous on `stream` (a cudaStream_t, may be NULL).  Photon g of batch b uses the
 * Philox4x32-10 stream keyed by (seed, batch_id, g), so any partition of photons or batches over
 * launches / GPUs gives bit-identical tallies.
 * Replaces the loop `for ray in scene.emit(n): photon_tracer.follow(scene, ray)` and
 * scene.simulate(num_rays, workers) (reference miniplant/simulation_runner.py:38-41,62-63).
 */
int lscpm_trace(LscpmScene* scene, const LscpmBatchDesc* batches, int32_t n_batches,
                const LscpmSpectraDesc* spectra, uint64_t photon_begin, uint64_t photon_count,
                uint64_t seed, int64_t* tallies_dev, void* stream);

/* Host buffers in, host tallies out: lowers the descriptors into the scene's pinned staging buffer, uploads them with
 * one copy, zeroes a tally block in the scene's device scratch, launches, copies the rows back into
 * tallies_host[n_batches*(n_bins+LSCPM_N_COUNTERS)] and synchronises (the scene's own stream; no allocation per call
 * once the scratch has grown to the job's size; calls on one scene are serialised by a mutex).  This is the call behind
 * _common_simulation_runner (reference miniplant/simulation_runner.py:65-66 reads count[REACT]/N) and behind the
 * drivers' one launch per year (full_simulation.py:64-106). */
int lscpm_trace_host(LscpmScene* scene, const LscpmBatchDesc* batches, int32_t n_batches,
                     const LscpmSpectraDesc* spectra, uint64_t photon_begin, uint64_t photon_count,
                     uint64_t seed, int64_t* tallies_host);

/* ---- resident batch plans --------------------------------------------------------------------- */
/* The drivers trace thousands of time points (reference miniplant/full_simulation.py:106 applies one
 * simulation per DataFrame row).  A plan lowers and uploads the batch descriptors and spectra once (ONE device
 * allocation, ONE copy from pinned memory), so that repeated launches (more photons, other seeds, other photon
 * ranges) start from inputs already resident in device memory.  lscpm_trace == plan_create + trace_plan +
 * plan_destroy.  `batches` may be any contiguous LscpmBatchDesc[n] - the Python binding passes a structured
 * NumPy array with this layout, filled column-wise. */
typedef struct LscpmPlan LscpmPlan;
int lscpm_plan_create(LscpmScene* scene, const LscpmBatchDesc* batches, int32_t n_batches,
                      const LscpmSpectraDesc* spectra, LscpmPlan** out);
int lscpm_plan_destroy(LscpmPlan* plan);
int lscpm_trace_plan(LscpmPlan* plan, uint64_t photon_begin, uint64_t photon_count, uint64_t seed,
                     int64_t* tallies_dev, void* stream);
/* Bytes lscpm_plan_create copied host->device for this plan (lowered batch descriptors + spectra). */
uint64_t lscpm_plan_h2d_bytes(const LscpmPlan* plan);
/* Number of kernel launches the library has issued on the calling process so far (all entry points). */
uint64_t lscpm_launch_count(void);

/* ---- deterministic per-ray entry points (parity check #1, and pvtrace next_hit) -------------- */
/*
 * For n rays given in WORLD coordinates (pos[3n], dir[3n], unit directions): the next interface
 * exactly as pvtrace's next_hit reports it (reference call site miniplant/simulation_runner.py:45-48)
 * plus the deterministic part of the surface interaction (container = nearest node the ray leaves
 * exactly once; adjacent = the hit node, or the container of what lies beyond when the ray leaves its
 * own container; tied intersections in scene-graph pre-order).  All arrays are host memory.
 *   hit/container/adjacent[n]  node indices (-1: none; hit = -1 means the ray leaves the scene)
 *   face[n]                    face of the hit node (box: -x,+x,-y,+y,-z,+z = 0..5; cylinder: side,-cap,+cap)
 *   distance[n], point[3n], normal[3n] (facing along the ray), reflectivity[n],
 *   reflected[3n], refracted[3n] (zero vector under total internal reflection)
 */
int lscpm_probe(LscpmScene* scene, int64_t n, const double* pos, const double* dir,
                int32_t* hit, int32_t* container, int32_t* adjacent, int32_t* face,
                double* distance, double* point, double* normal, double* reflectivity,
                double* reflected, double* refracted);

/* First `n` rays of a batch as the device generates them (world frame) — replaces scene.emit(n)
 * (reference miniplant/simulation_runner.py:38).  pos/dir are [3n], wavelength [n]; host memory. */
int lscpm_emit(LscpmScene* scene, const LscpmBatchDesc* batch, const LscpmSpectraDesc* spectra,
               uint64_t photon_begin, int64_t n, uint64_t seed, double* pos, double* dir, double* wavelength);

/*
 * Follow n caller-supplied rays (world frame) to termination, recording up to max_steps steps each:
 * replaces photon_tracer.follow(scene, ray) returning the (ray, event) history
 * (reference miniplant/simulation_runner.py:39-43).  Photon i uses RNG stream (seed, batch_id, photon_begin+i).
 *   n_steps[n]; step_event[n*max_steps]; step_node[n*max_steps]; step_pos[3*n*max_steps];
 *   step_dir[3*n*max_steps]; step_wavelength[n*max_steps]; final_bin[n]
 */
int lscpm_follow(LscpmScene* scene, int64_t n, const double* pos, const double* dir, const double* wavelength,
                 uint32_t batch_id, uint64_t photon_begin, uint64_t seed, int32_t max_steps,
                 int32_t* n_steps, int32_t* step_event, int32_t* step_node, double* step_pos, double* step_dir,
                 double* step_wavelength, int32_t* final_bin);

#ifdef __cplusplus
}
#endif
#endif /* LSCPM_H */
