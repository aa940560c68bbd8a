// lscpm_pool.cuh — K2p: the persistent photon kernel with a warp-private photon pool (phase-major execution).
//
// Why (profiles/r1_final_*): the lane-per-photon kernel K2 issues 84 % of its warp slots but only 14.7 of the
// 32 lanes of an issued instruction do work, because the 32 photons of a warp want different things at every point of an
// iteration (plate flight | tube geometry, then Fresnel | absorption, then generation | re-emission) and the warp runs all
// of them one after the other.  A Markov model of the photon's phases reproduces that figure for 32 photons per warp (14.4)
// and says what changes it: MORE PHOTONS THAN LANES per scheduling unit — 64 per warp give 26 lanes, 96 give 31.
//
// So a warp owns NSLOT (96) photon records in shared memory instead of 32 in registers.  Every record waits in exactly one
// phase list; a warp iteration picks the fullest list, pops up to 32 records of it, runs THAT phase's code for them (all
// lanes in the same code), stores what the phase changed and pushes each record onto the list of its next phase.  No
// barrier, no atomics on the lists (one warp, warp-uniform counters), no hand-over between warps: what made the block-level
// regrouping experiments of round 1 lose (K2s: the block waits for its slowest warp; page queues: queue management cost as
// much as the photon work) does not exist here.  Cost: four 16-byte shared-memory loads / three stores per phase
// execution and ~50 instructions of list bookkeeping (one packed counter word, one MATCH + one REDUX per push).
//
// What the first version's profile asked for (profiles/r2_pool_v1_*): it executed 25 % fewer warp instructions than K2 at
// 25.0 active lanes but issued only 59 % of its slots - 18 % of the stall samples were instruction-cache misses (52 KB of
// SASS against a 32 KB L1.5; warps in different phases fetch different code), 17 % waits on global loads (219 KB of
// shared memory left 28 KB of L1 for the spectra, hit rate 83 %).  Hence the 64-byte record (the attenuation
// coefficients are looked up from the wavelength when a flight needs them instead of travelling), at most 196 KB of
// shared memory, and compact bookkeeping.
//
//   PH_FRESH  empty record: takes the next photon index of the warp's work item, generation draw, entry Fresnel test,
//             and the photon's first flight (it is in registers and every lane wants the same thing)
//   PH_PLATE  photon in the plastic (scenes with plate flights): step draw + plate_flight, then the interaction the
//             flight ends in - absorption (component pick, quantum yield) or the Fresnel test on the interface reached
//   PH_TUBE   photon in a tube wall / the mixture (every medium when the scene has no plate flights): step draw +
//             compute_hit + free path + move, then the interaction as above
//   (PH_SURF  a Fresnel-test phase of its own existed up to v4; the test now runs inside the flight's execution)
//   PH_EMIT   photon absorbed by the luminophore, to be re-emitted: emission draw, and the flight that follows
//
// The phase code is the very device code of K2 / K2s (advance_hit, interact_*, source_*) fed with the same Philox words in
// the same order, so the tallies are bit-identical to K2's for every scene, batch shape and photon offset
// (tests/test_gpu_sorted_kernel.py covers all three kernels).
//
// EXT (scenes with boxes outside the plate - the photovoltaic variants): the air around the plate is a medium with geometry
// of its own (outside_step), and photons travelling in it wait in a list of their own: the PH_PLATE list when the scene has
// no plate flights, a fifth list (PH_AIR, its length in a second counter word) when it has.  The surface bookkeeping
// word (`aux`) that the EXIT classification of these scenes needs travels in the record's spare word (G3.z).
#pragma once
#include "lscpm_device.cuh"

namespace lscpm {

constexpr int PH_FRESH = 0, PH_PLATE = 1, PH_TUBE = 2, PH_EMIT = 3, N_PH = 4;
// photons in the air around the plate (scenes with outside boxes): the plate flights' list when the scene has none, a
// fifth list with a counter of its own otherwise
constexpr int PH_AIR = 4;
__host__ __device__ constexpr int pool_out_phase(bool flight) { return flight ? PH_AIR : PH_PLATE; }
__host__ __device__ constexpr int pool_lists(bool ext, bool flight) { return (ext && flight) ? 5 : N_PH; }

// record = 4 x uint4 (SoA over the slots of a warp):
//   G0: xr, y, z, kc | medium << 12 | steps << 16
//   G1: dx, dy, dz, wl
//   G2: photon index lo, hi, batch (index into the plan), batch id (Philox counter word)
//   G3: event counters, emission table of a pending re-emission (PH_EMIT), aux (scenes with outside boxes only: surface
//       bookkeeping of the EXIT classification; it carries the emission table itself there), Philox call counter
constexpr int POOL_GROUPS = 4;

__host__ __device__ constexpr int pool_list_u4(int nslot, int n_lists = N_PH) { return (n_lists * nslot + 15) / 16; }
__host__ __device__ constexpr size_t pool_smem_bytes(int warps, int nslot, int stride, int n_lists = N_PH) {
    return (size_t)warps * (size_t)(POOL_GROUPS * nslot + pool_list_u4(nslot, n_lists) + (stride + 3) / 4) * 16u;
}

template <bool FLIGHT, bool EXT = false>
__device__ __forceinline__ int fly_phase(int medium) {
    if (EXT && medium == MED_OUT) return pool_out_phase(FLIGHT);
    return (FLIGHT && medium == MED_PLATE) ? PH_PLATE : PH_TUBE;
}

// List lengths, warp-uniform, packed into ONE word: phase k in byte k (NSLOT <= 255).
//
// Push the records the active lanes hold onto the lists of their next phases: lanes with the same target find each other
// with one MATCH and rank themselves inside the group; the group leaders' sizes are summed into the packed counters with
// one REDUX - independent of how many different targets there are.
// FIVE: the fifth list (PH_AIR) has its length in `cnt_air`; its group is counted with one ballot.
template <int NSLOT, bool FIVE>
__device__ __forceinline__ void pool_push(unsigned char* __restrict__ lists, unsigned& cnt, unsigned& cnt_air, bool act, int nph, int slot, unsigned lane_lt) {
    const unsigned grp = __match_any_sync(0xffffffffu, act ? nph : 7);   // idle lanes form a group of their own that is not counted
    const int rank = __popc(grp & lane_lt);
    const bool air = FIVE && nph == PH_AIR;
    const unsigned shift = air ? 0u : 8u * (unsigned)nph;
    const unsigned base = air ? cnt_air : ((cnt >> shift) & 0xffu);
    if (act) lists[nph * NSLOT + (int)base + rank] = (unsigned char)slot;
    const unsigned contrib = (act && rank == 0 && !air) ? (unsigned)__popc(grp) << shift : 0u;
    cnt += __reduce_add_sync(0xffffffffu, contrib);
    if (FIVE) cnt_air += (unsigned)__popc(__ballot_sync(0xffffffffu, act && air));
}

__device__ __forceinline__ void pool_tally(unsigned long long* __restrict__ tallies, unsigned int* hist, int n_bins, int stride,
                                           int hbatch, int batch, int bin, unsigned n_events) {
    LSCPM_CHECK(bin >= 0 && bin < n_bins);
    if (batch == hbatch) {
        atomicAdd(hist + bin, 1u);
        if (n_events & 0xffffu) atomicAdd(hist + n_bins, n_events & 0xffffu);
        if (n_events >> 16) atomicAdd(hist + n_bins + 1, n_events >> 16);
    } else {   // straggler of a batch the warp has moved on from
        unsigned long long* row = tallies + (size_t)batch * stride;
        atomicAdd(row + bin, 1ULL);
        if (n_events & 0xffffu) atomicAdd(row + n_bins, (unsigned long long)(n_events & 0xffffu));
        if (n_events >> 16) atomicAdd(row + n_bins + 1, (unsigned long long)(n_events >> 16));
    }
}

}  // namespace lscpm

// The kernel itself is compiled inside lscpm.cu, after TraceArgs / flush_row (this header is not stand-alone).
#ifdef LSCPM_POOL_KERNEL_BODY

template <bool EXT, bool FLIGHT, int WARPS, int NSLOT>
__global__ void __launch_bounds__(WARPS * 32, 1)
trace_pool_kernel(const __grid_constant__ lscpm::DevScene S, const __grid_constant__ TraceArgs A) {
    using namespace lscpm;
    static_assert(NSLOT % 8 == 0 && NSLOT <= 255, "slot numbers and list lengths travel as bytes");
    constexpr bool FIVE = EXT && FLIGHT;
    constexpr int N_LISTS = pool_lists(EXT, FLIGHT), PH_OUT = pool_out_phase(FLIGHT);
    extern __shared__ uint4 pool_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lane_lt = (1u << lane) - 1u;
    const int stride = S.n_bins + LSCPM_N_COUNTERS;
    const int per_warp = POOL_GROUPS * NSLOT + pool_list_u4(NSLOT, N_LISTS) + (stride + 3) / 4;
    uint4* G = pool_smem + (size_t)warp * per_warp;
    unsigned char* lists = reinterpret_cast<unsigned char*>(G + POOL_GROUPS * NSLOT);
    unsigned int* hist = reinterpret_cast<unsigned int*>(G + POOL_GROUPS * NSLOT + pool_list_u4(NSLOT, N_LISTS));
    for (int b = lane; b < stride; b += 32) hist[b] = 0u;
    for (int i = lane; i < NSLOT; i += 32) lists[PH_FRESH * NSLOT + i] = (unsigned char)i;
    __syncwarp();

    unsigned cnt = (unsigned)NSLOT;         // packed list lengths: every record starts empty (PH_FRESH = byte 0)
    unsigned cnt_air = 0u;                  // length of the fifth list (FIVE)
    int hbatch = -1;                        // batch whose photons the shared row accumulates
    unsigned long long chunk_base = 0;      // work item: photons chunk_base + [c_next, c_end) of batch `cbatch` are unclaimed
    unsigned c_next = 0, c_end = 0;
    int cbatch = 0;
    bool more = true;
    unsigned items_since_flush = 0;

    for (;;) {
        // ---- the fullest list whose phase can run: max over (length << 3 | phase) ---------------------------------
        const unsigned fresh_len = (c_next < c_end || more) ? (cnt & 0xffu) : 0u;
        unsigned key = fresh_len << 3;
        key = max(key, (__byte_perm(cnt, 0u, 0x4441) << 3) | (unsigned)PH_PLATE);
        key = max(key, (__byte_perm(cnt, 0u, 0x4442) << 3) | (unsigned)PH_TUBE);
        key = max(key, ((cnt >> 24) << 3) | (unsigned)PH_EMIT);
        if (FIVE) key = max(key, (cnt_air << 3) | (unsigned)PH_AIR);
        const int best = (int)(key >> 3);
        if (best == 0) break;
        const int ph = (int)(key & 7u);

        if (ph == PH_FRESH && c_next >= c_end) {   // claim the next work item
            unsigned long long id = 0;
            if (lane == 0) id = atomicAdd(A.work_counter, 1ULL);
            id = __shfl_sync(FULL_MASK, id, 0);
            if (id >= A.n_chunks) { more = false; continue; }
            cbatch = (int)(id / A.chunks_per_batch);
            const unsigned long long first = (id % A.chunks_per_batch) * (unsigned long long)A.chunk;
            chunk_base = A.photon_begin + first;
            c_next = 0;
            c_end = (unsigned)min((unsigned long long)A.chunk, A.photon_count - first);
            // the 32-bit shared row is flushed when the batch changes and, within one batch, often enough not to wrap
            if (cbatch != hbatch || ++items_since_flush >= 256u) {
                if (hbatch >= 0) flush_row(hist, A.tallies + (size_t)hbatch * stride, stride, lane);
                hbatch = cbatch;
                items_since_flush = 0;
            }
        }

        // ---- pop up to 32 records of that list ---------------------------------------------------------------
        int have = best;
        if (ph == PH_FRESH) have = min(have, (int)(c_end - c_next));
        const int n = min(32, have);
        const bool act = lane < n;
        int slot = 0;
        if (act) slot = lists[ph * NSLOT + best - n + lane];
        LSCPM_CHECK(slot >= 0 && slot < NSLOT);
        if (FIVE && ph == PH_AIR) cnt_air -= (unsigned)n;
        else cnt -= (unsigned)n << (8 * ph);
        int nph = PH_FRESH;
        int tbin = -1, tbatch = 0;          // a photon that ended in this execution: tallied once, below
        unsigned tevents = 0u;

        {
            // ---- source (PH_FRESH: a new photon, PH_EMIT: a re-emission), then - in the SAME execution, with the photon
            // ---- still in registers - its first flight; or the flight of a record waiting in PH_PLATE / PH_TUBE.
            // A photon that has just been generated or re-emitted always flies next, and all lanes of a source execution
            // do (but the 4 % reflected at first contact), so running the flight here costs no lanes and saves one phase
            // execution per generation / emission: its bookkeeping and a whole record round trip through shared memory.
            const bool is_source = ph == PH_FRESH || ph == PH_EMIT;      // warp-uniform
            const bool fresh = ph == PH_FRESH;
            bool fly = act && !is_source;
            Photon p;
            p.a_plate = p.a_outer = p.a_inner = 0.f; p.aux = 0;
            Rng rng; rng.key = A.key;
            int batch = 0;
            unsigned n_events = 0u;
            if (is_source) {
                if (act) {
                    if (fresh) {
                        const unsigned long long g = chunk_base + c_next + (unsigned)lane;
                        batch = cbatch;
                        rng.p_lo = (uint32_t)g; rng.p_hi = (uint32_t)(g >> 32); rng.batch = A.batches[batch].batch_id; rng.calls = 0;
                        p.xr = p.y = p.z = p.wl = 0.f; p.kc = p.medium = p.steps = 0;
                    } else {
                        const uint4 g0 = G[0 * NSLOT + slot], g2 = G[2 * NSLOT + slot], g3 = G[3 * NSLOT + slot];
                        p.xr = __uint_as_float(g0.x); p.y = __uint_as_float(g0.y); p.z = __uint_as_float(g0.z);
                        p.kc = g0.w & 0xfff; p.medium = (g0.w >> 12) & 15; p.steps = g0.w >> 16;
                        p.wl = __uint_as_float(G[1 * NSLOT + slot].w);
                        p.aux = EXT ? (int)g3.z : (int)(g3.y << 16);    // carries the emission table (bits 16-23)
                        rng.p_lo = g2.x; rng.p_hi = g2.y; batch = (int)g2.z; rng.batch = g2.w; rng.calls = g3.w;
                        n_events = g3.x;
                    }
                    const uint4 v = rng.next(A.rk);
                    WavelengthDraw wd;
                    int bin = -1;
                    if (fresh) bin = source_fresh<EXT, true>(S, A.batches[batch], A.spec_x, A.spec_cdf, A.spec_guide, rng, v, p, wd, nullptr, n_events);
                    else source_emit(S, v, p, wd);
                    source_wavelength(p, wd);
                    if (bin >= 0) {   // reflected off the plate at first contact, or never touched it: the record stays empty
                        tbin = bin; tbatch = batch; tevents = n_events;
                    } else {
                        fly = true;
                    }
                }
                if (fresh) c_next += (unsigned)n;
            }
            // A new photon of a scene with outside boxes may start in the air, on a plate face or inside the plate.  When
            // every lane's photon starts in the air (the reference's lights do: they sit 1 m up-beam of the plate), the first
            // step in the air follows right here; otherwise there is no common next step and the photons wait in the lists
            // of their media.
            bool fresh_outside = false;
            if (EXT && fresh) {
                fresh_outside = __all_sync(FULL_MASK, !fly || p.medium == MED_OUT);
                if (!fresh_outside && fly) {
                    nph = fly_phase<FLIGHT, EXT>(p.medium);
                    G[0 * NSLOT + slot] = make_uint4(__float_as_uint(p.xr), __float_as_uint(p.y), __float_as_uint(p.z),
                                                     (unsigned)p.kc | ((unsigned)p.medium << 12) | ((unsigned)p.steps << 16));
                    G[1 * NSLOT + slot] = make_uint4(__float_as_uint(p.dx), __float_as_uint(p.dy), __float_as_uint(p.dz), __float_as_uint(p.wl));
                    G[2 * NSLOT + slot] = make_uint4(rng.p_lo, rng.p_hi, (unsigned)batch, rng.batch);
                    G[3 * NSLOT + slot] = make_uint4(n_events, 0u, (unsigned)p.aux, rng.calls);
                    fly = false;
                }
            }
            // ---- step draw, nearest interface, free path, move; absorption is settled here as well ---------------
            bool keep = false;              // the photon lives on: its record is written back below, by all such lanes at once
            unsigned pend = 0u;             // emission table of a pending re-emission
            bool emitted = false;
            if (fly) {
                if (!is_source) {
                    const uint4 g0 = G[0 * NSLOT + slot], g1 = G[1 * NSLOT + slot], g2 = G[2 * NSLOT + slot], g3 = G[3 * NSLOT + slot];
                    p.xr = __uint_as_float(g0.x); p.y = __uint_as_float(g0.y); p.z = __uint_as_float(g0.z);
                    p.kc = g0.w & 0xfff; p.medium = (g0.w >> 12) & 15; p.steps = g0.w >> 16;
                    p.dx = __uint_as_float(g1.x); p.dy = __uint_as_float(g1.y); p.dz = __uint_as_float(g1.z); p.wl = __uint_as_float(g1.w);
                    rng.p_lo = g2.x; rng.p_hi = g2.y; batch = (int)g2.z; rng.batch = g2.w; rng.calls = g3.w;
                    n_events = g3.x;
                    if (EXT) p.aux = (int)g3.z;
                }
                const uint4 w = rng.next(A.rk);
                Hit h; int cls = CLS_CELL, extra = 0; float alpha = 0.f;
                int bin = -1;
                const bool outside = EXT && (ph == PH_OUT || fresh_outside);   // warp-uniform: only photons in the air
                if (outside) {
                    // one follow() iteration in the air around the plate: next hit among the plate and the outside boxes
                    OutState q;
                    q.x = abs_x(S, p); q.y = p.y; q.z = p.z + S.zc; q.dx = p.dx; q.dy = p.dy; q.dz = p.dz; q.aux = p.aux; q.steps = p.steps;
                    const OutResult r = outside_step(S, q, u01(w.y));
                    p.dx = r.dx; p.dy = r.dy; p.dz = r.dz; p.aux = r.aux; p.steps = r.steps;
                    if (counts_as_event(r.event)) n_events += 1u + (unsigned)r.extra;
                    if (r.entered) { p.y = r.y; locate_in_plate(S, p, r.x, r.z); }
                    else { p.kc = 0; p.xr = r.x - axis_x(S, 0); p.y = r.y; p.z = r.z - S.zc; }
                    bin = r.bin;
                } else {
                    bin = advance_hit<true, FLIGHT, true, true>(S, p, w.x, h, cls, extra, alpha);
                    if (FIVE) p.aux = note_folded(p.aux, extra);
                    n_events += (unsigned)extra;
                }
                if (bin < 0 && !outside) {
                    if (cls == CLS_ABSORB) {
                        n_events += 1u;
                        int ev;
                        bin = interact_absorb_alpha(S, p, h.container, alpha, u01(w.y), w.z, ev);
                        if (bin < 0) {   // re-emitted by the luminophore: the emission table travels in the pending word
                            n_events += 0x10000u;
                            emitted = true;
                            pend = (unsigned)(p.aux >> 16) & 0xffu;
                        }
                    } else if (cls != CLS_CELL) {
                        // Fresnel test on the interface the photon was moved to, in the same execution: about half of a
                        // flight's lanes are here, and a phase of its own would cost them more in bookkeeping and record
                        // traffic than the idle lanes cost here
                        n_events += 1u;
                        int ev;
                        bin = interact_surface<EXT>(S, p, h.kind, h.face, h.container, h.adj, u01(w.y), ev);
                    }
                }
                if (bin >= 0) { tbin = bin; tbatch = batch; tevents = n_events; }
                else keep = true;
            }
            // the outcomes above diverge (absorbed | re-emitted | cell crossing | surface); without this convergence point the
            // compiler duplicates the write-back into every one of them (profile of v4: the stores ran at 12.5 lanes)
            __syncwarp();
            if (keep) {
                // the next phase is chosen here, once, by all surviving lanes (in the branches above it ran at 6 lanes)
                nph = emitted ? PH_EMIT : fly_phase<FLIGHT, EXT>(p.medium);
                G[0 * NSLOT + slot] = make_uint4(__float_as_uint(p.xr), __float_as_uint(p.y), __float_as_uint(p.z),
                                                 (unsigned)p.kc | ((unsigned)p.medium << 12) | ((unsigned)p.steps << 16));
                G[1 * NSLOT + slot] = make_uint4(__float_as_uint(p.dx), __float_as_uint(p.dy), __float_as_uint(p.dz), __float_as_uint(p.wl));
                if (fresh) G[2 * NSLOT + slot] = make_uint4(rng.p_lo, rng.p_hi, (unsigned)batch, rng.batch);
                G[3 * NSLOT + slot] = make_uint4(n_events, pend, EXT ? (unsigned)p.aux : 0u, rng.calls);
            }
        }
        if (tbin >= 0) pool_tally(A.tallies, hist, S.n_bins, stride, hbatch, tbatch, tbin, tevents);
        pool_push<NSLOT, FIVE>(lists, cnt, cnt_air, act, nph, slot, lane_lt);
        __syncwarp();   // records and lists written by one lane are read by another in a later iteration
    }
    if (hbatch >= 0) flush_row(hist, A.tallies + (size_t)hbatch * stride, stride, lane);
}

#endif  // LSCPM_POOL_KERNEL_BODY
