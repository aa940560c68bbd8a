// lscpm.cu — kernels and C-ABI of liblscpm.so (see include/lscpm.h for the contract and the reference
// interfaces each entry point replaces).  sm_100a only; no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lscpm.h"
#include "lscpm_device.cuh"

using namespace lscpm;

// =================================================================================================
// kernels
// =================================================================================================
// lane kernel block sizes: one 1024-thread block per SM is 4 % faster than eight 128-thread blocks on every workload
// measured (same registers, same occupancy; profiles/r1_ab_block_size.txt); the small block remains for scenes whose
// per-warp tally rows would not fit 32 times into shared memory
constexpr int TRACE_BLOCK_BIG = 1024, TRACE_BLOCK_SMALL = 128;
constexpr size_t TRACE_SMEM_BIG_LIMIT = 96 * 1024;
constexpr unsigned FULL_MASK = 0xffffffffu;

struct TraceArgs {
    const DevBatch* batches;
    const float* spec_x;
    const float* spec_cdf;
    const unsigned short* spec_guide;
    unsigned long long photon_begin;    // first photon index of every batch in this launch
    unsigned long long photon_count;    // photons per batch in this launch
    unsigned long long chunks_per_batch;
    unsigned long long n_chunks;
    unsigned int chunk;                  // photons per work item
    uint2 key;
    RoundKeys rk;                        // Philox round keys of `key` (constant-bank operands of the trace kernels)
    unsigned long long* work_counter;
    unsigned long long* tallies;         // [n_batches][n_bins + LSCPM_N_COUNTERS]
};

__device__ __forceinline__ bool counts_as_event(int ev) {
    return ev == EV_REFLECT || ev == EV_TRANSMIT || ev == EV_ABSORB || ev == EV_REACT || ev == EV_EMIT;
}

// K2: persistent photon kernel.  Every lane carries one photon in registers and steps it until it
// terminates; dead lanes are refilled at the top of the loop from the warp's current work item
// (ballot + popc ranks), work items are claimed with one atomic per warp.  A refilled lane generates its
// photon from the SAME Philox draw the other lanes use for their step and then joins the common step for
// its entry Fresnel test.  The photon's result depends only on (seed, batch id, photon index), so the
// scheduling does not influence the tallies.
// Per-warp tally rows live in shared memory: a warp adds the photons of the batch it is currently drawing from
// to its own row with shared-memory atomics and flushes the row to the global tallies (one 64-bit atomic per
// non-zero bin) when it moves on to another batch and at the end.  Stragglers of an earlier batch go to global
// memory directly.  Without this the per-batch counters serialise in L2 (one hot address per batch).
__device__ __forceinline__ void flush_row(unsigned int* hist, unsigned long long* row, int stride, int lane) {
    __syncwarp();
#pragma unroll 1
    for (int b = lane; b < stride; b += 32) {
        const unsigned int v = hist[b];
        if (v) { atomicAdd(row + b, (unsigned long long)v); hist[b] = 0u; }
    }
    __syncwarp();
}

// EXT: the scene has boxes outside the plate (photovoltaic variants); the standard scene runs an EXT = false
// instantiation, which contains none of that code.  FLIGHT: photons in the plastic go straight to their next real event
// (plate_flight); chosen when the plate holds a luminophore, i.e. when there is trapped light to fold.
template <bool EXT, bool FLIGHT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK)
trace_kernel(const __grid_constant__ DevScene S, const __grid_constant__ TraceArgs A) {
    extern __shared__ unsigned int smem_rows[];
    const int lane = threadIdx.x & 31;
    const unsigned lane_lt = (1u << lane) - 1u;
    const int stride = S.n_bins + LSCPM_N_COUNTERS;
    unsigned int* hist = smem_rows + (threadIdx.x >> 5) * stride;
    for (int b = lane; b < stride; b += 32) hist[b] = 0u;
    __syncwarp();
    int hbatch = -1;                        // warp-uniform: batch whose photons the shared row accumulates

    // warp-uniform work item: photons chunk_base + [c_next, c_end) of batch `cbatch` are unclaimed
    unsigned long long chunk_base = 0;
    unsigned c_next = 0, c_end = 0;
    int cbatch = 0;
    bool more = true;                       // work items may remain in the global queue
    unsigned items_since_flush = 0u;

    Photon p;
    Rng rng;
    int batch = 0;
    bool alive = false;
    unsigned n_events = 0;                  // low 16 bits: interaction events, high 16 bits: emissions

    for (;;) {
        // ---- step every live photon ---------------------------------------------------------------
        int status = -1;
        if (alive) {
            const uint4 w = rng.next(A.rk);
            StepOut so;
            status = photon_step<EXT, FLIGHT>(S, p, w, so);
            n_events += (unsigned)so.extra + (counts_as_event(so.event) ? 1u : 0u);
            if (so.event == EV_EMIT) n_events += 0x10000u;
        }
        // ---- tally the photons that ended -----------------------------------------------------------
        if (status >= 0) {
            LSCPM_CHECK(status < S.n_bins);
            if (batch == hbatch) {
                atomicAdd(hist + status, 1u);
                atomicAdd(hist + S.n_bins, n_events & 0xffffu);
                if (n_events >> 16) atomicAdd(hist + S.n_bins + 1, n_events >> 16);
            } else {
                unsigned long long* row = A.tallies + (size_t)batch * stride;
                atomicAdd(row + status, 1ULL);
                atomicAdd(row + S.n_bins, (unsigned long long)(n_events & 0xffffu));
                if (n_events >> 16) atomicAdd(row + S.n_bins + 1, (unsigned long long)(n_events >> 16));
            }
            alive = false;
        }
        // ---- refill dead lanes from the warp's work item --------------------------------------------
        bool want = !alive;
        bool fresh = false;
        while (true) {
            const unsigned pending = __ballot_sync(FULL_MASK, want);
            if (!pending) break;
            if (c_next >= c_end) {
                if (!more) break;
                unsigned long long id = 0;
                if (lane == 0) id = atomicAdd(A.work_counter, 1ULL);
                id = __shfl_sync(FULL_MASK, id, 0);
                if (id >= A.n_chunks) { more = false; break; }
                cbatch = (int)(id / A.chunks_per_batch);
                const unsigned long long first = (id % A.chunks_per_batch) * (unsigned long long)A.chunk;
                chunk_base = A.photon_begin + first;
                c_next = 0;
                c_end = (unsigned)min((unsigned long long)A.chunk, A.photon_count - first);
                // the 32-bit shared row is flushed when the batch changes and, within one batch, every 256 work items
                // (<= 1M photons): a single huge batch cannot wrap its event counters
                if (cbatch != hbatch || ++items_since_flush >= 256u) {
                    if (hbatch >= 0) flush_row(hist, A.tallies + (size_t)hbatch * stride, stride, lane);
                    hbatch = cbatch;
                    items_since_flush = 0u;
                }
            }
            const unsigned avail = c_end - c_next;
            const unsigned rank = __popc(pending & lane_lt);
            if (want && rank < avail) {
                const unsigned long long g = chunk_base + c_next + rank;
                batch = cbatch;
                rng.key = A.key; rng.p_lo = (uint32_t)g; rng.p_hi = (uint32_t)(g >> 32);
                rng.batch = A.batches[batch].batch_id; rng.calls = 0;
                want = false;
                fresh = true;
            }
            c_next += min((unsigned)__popc(pending), avail);
        }
        // ---- source block: new photons and re-emissions ---------------------------------------------
        if (fresh || status == STEP_EMIT) {
            const uint4 v = rng.next(A.rk);
            WavelengthDraw wd;
            int bin = -1;
            if (fresh) {
                bin = source_fresh<EXT, true>(S, A.batches[batch], A.spec_x, A.spec_cdf, A.spec_guide, rng, v, p, wd, nullptr, n_events);
                alive = true;
            } else {
                source_emit(S, v, p, wd);
            }
            source_finish(S, p, wd);
            if (bin >= 0) {   // reflected off the plate at first contact, or never touched it
                if (batch == hbatch) {
                    atomicAdd(hist + bin, 1u);
                    if (n_events) atomicAdd(hist + S.n_bins, n_events);
                } else {   // the warp moved on to another work item within this refill (work items shorter than a warp)
                    unsigned long long* row = A.tallies + (size_t)batch * stride;
                    atomicAdd(row + bin, 1ULL);
                    if (n_events) atomicAdd(row + S.n_bins, (unsigned long long)n_events);
                }
                alive = false;
            }
        }
        if (!__any_sync(FULL_MASK, alive) && !more && c_next >= c_end) break;
    }
    if (hbatch >= 0) flush_row(hist, A.tallies + (size_t)hbatch * stride, stride, lane);
}

// K2s: block-sorted persistent photon kernel.  Photons in a warp stop doing the same thing after the common half of
// an iteration (advance): some meet a flat face, some a tube wall, some are absorbed, some need a new photon.  Left in
// place they serialise (the lane-per-photon kernel above averages 15.6 of 32 active lanes).  Here every iteration the
// whole block re-deals its photons through shared memory so that each warp holds (mostly) one class of pending work:
// a counting sort over N_CLS classes (warp match -> per-warp 16-bit counts -> one packed 2-entries-per-lane scan that
// every warp repeats for itself), a scatter of the 20-word photon record to its sorted slot, a linear read back.
// New photon indices go to the FRESH class by rank from a double-buffered block-level work cursor (current + prefetched
// work item, refilled by thread 0 with one global atomic per item); the tally of a photon that ended is made by the
// FRESH class as well, warp-aggregated, into one shared-memory row per block for the batch of the current work item.
#ifndef SORT_BLOCK
#define SORT_BLOCK 256
#endif
constexpr int SORT_THREADS = SORT_BLOCK;
#ifndef SORT_MIN_BLOCKS
#define SORT_MIN_BLOCKS (1024 / SORT_BLOCK)
#endif

#ifdef SORT_TIMING   // diagnostic build: where do the warps of a block spend an iteration?
__device__ unsigned long long g_sort_timing[64];
#endif

struct WorkCursor {
    unsigned long long cur_base, nxt_base;
    unsigned cur_left, nxt_left;
    int cur_batch, nxt_batch;
    int more, pad;
};

template <int BLOCK>
struct SortShared {
    uint4 x[5][BLOCK];
    unsigned short counts[64];
    WorkCursor cursor[2];
};

__device__ __forceinline__ void claim_item(const TraceArgs& A, unsigned long long& base, unsigned& left, int& batch, int& more) {
    left = 0;
    if (!more) return;
    const unsigned long long id = atomicAdd(A.work_counter, 1ULL);
    if (id >= A.n_chunks) { more = 0; return; }
    batch = (int)(id / A.chunks_per_batch);
    const unsigned long long first = (id % A.chunks_per_batch) * (unsigned long long)A.chunk;
    base = A.photon_begin + first;
    left = (unsigned)min((unsigned long long)A.chunk, A.photon_count - first);
}

template <bool EXT, bool FLIGHT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, SORT_MIN_BLOCKS)
trace_sorted_kernel(const __grid_constant__ DevScene S, const __grid_constant__ TraceArgs A) {
    constexpr int NW = BLOCK / 32;
    static_assert(N_CLS * NW <= 64, "class table must fit two 16-bit entries per lane");
    __shared__ SortShared<BLOCK> sm;
    extern __shared__ unsigned int hist[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lane_lt = (1u << lane) - 1u;
    const int stride = S.n_bins + LSCPM_N_COUNTERS;
    for (int b = tid; b < stride; b += BLOCK) hist[b] = 0u;
    if (tid < 64) sm.counts[tid] = 0;
    if (tid == 0) {
        WorkCursor c;
        c.more = 1; c.cur_batch = c.nxt_batch = 0; c.cur_base = c.nxt_base = 0; c.pad = 0;
        claim_item(A, c.cur_base, c.cur_left, c.cur_batch, c.more);
        claim_item(A, c.nxt_base, c.nxt_left, c.nxt_batch, c.more);
        sm.cursor[0] = c;
    }
    __syncthreads();
    int hbatch = -1;                        // block-uniform: batch whose photons the shared row accumulates
    int it = 0;                             // iteration parity selects the cursor copy
    unsigned iterations = 0u;

    Photon p;
    p.xr = p.y = p.z = p.dx = p.dy = p.dz = p.wl = p.a_plate = p.a_outer = p.a_inner = 0.f;
    p.kc = p.medium = p.steps = p.aux = 0;
    Rng rng;
    rng.key = A.key; rng.p_lo = rng.p_hi = rng.batch = rng.calls = 0;
    int batch = 0;
    unsigned n_events = 0;                  // low 16 bits: interaction events, high 16 bits: emissions
    unsigned pend = CLS_FRESH;              // pending interaction (class in the low 4 bits)
    float u_surf = 0.f;
    uint32_t w_z = 0;
    int pbin = -1;                          // FRESH class: tally bin of the photon this lane has just finished

#ifdef SORT_TIMING
    long long t_work = 0, t_wait = 0, t_sort = 0, t_mark = clock64(), n_it = 0;
    long long t_cls[N_CLS] = {0}, n_cls[N_CLS] = {0}, t_mixed = 0, n_mixed = 0;   // g_sort_timing[8 + 2c], [9 + 2c]; mixed at [30], [31]
    int wcls = -1;
#endif
    for (;; it ^= 1) {
        int cls = pend & 15;
        // ---- counting sort of the block's photons by class ------------------------------------------
        const unsigned mine = __match_any_sync(FULL_MASK, cls);
#ifdef SORT_TIMING
        { const long long now = clock64(); const long long d = now - t_mark; t_work += d; t_mark = now; ++n_it;
          if (wcls >= 0) { t_cls[wcls] += d; ++n_cls[wcls]; } else if (wcls == -2) { t_mixed += d; ++n_mixed; } }
#endif
        if (lane < N_CLS) sm.counts[lane * NW + warp] = 0;
        __syncwarp();
        if ((mine & lane_lt) == 0u) sm.counts[cls * NW + warp] = (unsigned short)__popc(mine);
        __syncthreads();
#ifdef SORT_TIMING
        { const long long now = clock64(); t_wait += now - t_mark; t_mark = now; }
#endif
        // packed exclusive scan of the class-major table: lane L owns entries 2L (low half) and 2L+1 (high half)
        const unsigned word = reinterpret_cast<const unsigned*>(sm.counts)[lane];
        const unsigned lo = word & 0xffffu, pair = lo + (word >> 16);
        unsigned incl = pair;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        const unsigned pre = (incl - pair) | ((incl - pair + lo) << 16);
        const int entry = cls * NW + warp;
        const unsigned base_word = __shfl_sync(FULL_MASK, pre, entry >> 1);
        const int dest = (int)((entry & 1) ? base_word >> 16 : base_word & 0xffffu) + __popc(mine & lane_lt);
        constexpr int E_FRESH = CLS_FRESH * NW, E_AFTER = (CLS_FRESH + 1) * NW, E_IDLE = CLS_IDLE * NW;
        const unsigned w_fresh = __shfl_sync(FULL_MASK, pre, E_FRESH >> 1), w_after = __shfl_sync(FULL_MASK, pre, E_AFTER >> 1);
        const unsigned w_idle = __shfl_sync(FULL_MASK, pre, E_IDLE >> 1);
        const int fresh_begin = (int)((E_FRESH & 1) ? w_fresh >> 16 : w_fresh & 0xffffu);
        const int fresh_end = (int)((E_AFTER & 1) ? w_after >> 16 : w_after & 0xffffu);
        const int idle_begin = (int)((E_IDLE & 1) ? w_idle >> 16 : w_idle & 0xffffu);
        if (idle_begin == 0) break;                      // every lane of the block is out of work
        LSCPM_CHECK(dest >= 0 && dest < BLOCK);
        // ---- the shared tally row follows the batch of the current work item --------------------------
        const int target = sm.cursor[it].cur_batch;
        // also every 2^14 iterations (<= 4M photon-steps per block): a single huge batch cannot wrap the 32-bit row
        if (target != hbatch || (++iterations & 0x3fffu) == 0u) {
            if (hbatch >= 0) {
                unsigned long long* row = A.tallies + (size_t)hbatch * stride;
                for (int b = tid; b < stride; b += BLOCK) {
                    const unsigned int v = hist[b];
                    if (v) { atomicAdd(row + b, (unsigned long long)v); hist[b] = 0u; }
                }
            }
            hbatch = target;
        }
        // ---- re-deal --------------------------------------------------------------------------------
        sm.x[0][dest] = make_uint4(__float_as_uint(p.xr), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(p.dx));
        sm.x[1][dest] = make_uint4(__float_as_uint(p.dy), __float_as_uint(p.dz), __float_as_uint(p.wl), __float_as_uint(p.a_plate));
        sm.x[2][dest] = make_uint4(__float_as_uint(p.a_outer), __float_as_uint(p.a_inner),
                                   (unsigned)p.kc | ((unsigned)p.medium << 12) | ((unsigned)p.steps << 16), (unsigned)p.aux);
        sm.x[3][dest] = make_uint4(rng.p_lo, rng.p_hi, (unsigned)batch, rng.calls);
        sm.x[4][dest] = make_uint4(n_events, pend, cls == CLS_FRESH ? (unsigned)pbin : __float_as_uint(u_surf), w_z);
        __syncthreads();
        {
            const uint4 q0 = sm.x[0][tid], q1 = sm.x[1][tid], q2 = sm.x[2][tid], q3 = sm.x[3][tid], q4 = sm.x[4][tid];
            p.xr = __uint_as_float(q0.x); p.y = __uint_as_float(q0.y); p.z = __uint_as_float(q0.z); p.dx = __uint_as_float(q0.w);
            p.dy = __uint_as_float(q1.x); p.dz = __uint_as_float(q1.y); p.wl = __uint_as_float(q1.z); p.a_plate = __uint_as_float(q1.w);
            p.a_outer = __uint_as_float(q2.x); p.a_inner = __uint_as_float(q2.y);
            p.kc = q2.z & 0xfff; p.medium = (q2.z >> 12) & 15; p.steps = q2.z >> 16; p.aux = (int)q2.w;
            rng.p_lo = q3.x; rng.p_hi = q3.y; batch = (int)q3.z; rng.calls = q3.w;
            n_events = q4.x; pend = q4.y; u_surf = __uint_as_float(q4.z); w_z = q4.w;
            pbin = (int)q4.z;
        }
        cls = pend & 15;
#ifdef SORT_TIMING
        { const long long now = clock64(); t_sort += now - t_mark; t_mark = now;
          const unsigned same = __match_any_sync(FULL_MASK, cls); wcls = same == FULL_MASK ? cls : -2; }
#endif
        const WorkCursor& cur = sm.cursor[it];
        if (tid == 0) {
            // next iteration's cursor: consume the indices the FRESH class takes below, keep two work items in hand
            WorkCursor c = cur;
            unsigned n = (unsigned)(fresh_end - fresh_begin);
            unsigned take = min(n, c.cur_left); c.cur_base += take; c.cur_left -= take; n -= take;
            take = min(n, c.nxt_left); c.nxt_base += take; c.nxt_left -= take;
            while (c.cur_left == 0u && (c.nxt_left > 0u || c.more)) {
                c.cur_base = c.nxt_base; c.cur_left = c.nxt_left; c.cur_batch = c.nxt_batch;
                claim_item(A, c.nxt_base, c.nxt_left, c.nxt_batch, c.more);
            }
            sm.cursor[it ^ 1] = c;
        }
        // ---- tallies of the photons that ended, warp-aggregated (the FRESH class holds them) -----------
        {
            const bool have = cls == CLS_FRESH && pbin >= 0;
            if (__any_sync(FULL_MASK, have)) {
                LSCPM_CHECK(!have || pbin < S.n_bins);
                const bool local = have && batch == hbatch;
                const unsigned grp = __match_any_sync(FULL_MASK, local ? pbin : -1);
                if (local && (grp & lane_lt) == 0u) atomicAdd(hist + pbin, (unsigned)__popc(grp));
                const unsigned ev = __reduce_add_sync(FULL_MASK, local ? (n_events & 0xffffu) : 0u);
                const unsigned em = __reduce_add_sync(FULL_MASK, local ? (n_events >> 16) : 0u);
                if (lane == 0) {
                    if (ev) atomicAdd(hist + S.n_bins, ev);
                    if (em) atomicAdd(hist + S.n_bins + 1, em);
                }
                if (have && !local) {
                    unsigned long long* row = A.tallies + (size_t)batch * stride;
                    atomicAdd(row + pbin, 1ULL);
                    atomicAdd(row + S.n_bins, (unsigned long long)(n_events & 0xffffu));
                    if (n_events >> 16) atomicAdd(row + S.n_bins + 1, (unsigned long long)(n_events >> 16));
                }
            }
        }
        // ---- the pending interaction, one class per warp (mostly) ------------------------------------
        int bin = -1;
        bool fly = false, source = false, fresh = false;
        if (cls == CLS_FRESH) {
            pbin = -1;
            const unsigned r = (unsigned)(tid - fresh_begin);
            unsigned long long g = 0; int nb = -1;
            if (r < cur.cur_left) { g = cur.cur_base + r; nb = cur.cur_batch; }
            else if (r - cur.cur_left < cur.nxt_left) { g = cur.nxt_base + (r - cur.cur_left); nb = cur.nxt_batch; }
            if (nb >= 0) {
                batch = nb;
                rng.p_lo = (uint32_t)g; rng.p_hi = (uint32_t)(g >> 32); rng.calls = 0;
                n_events = 0;
                source = fresh = true;
            } else if (!cur.more && cur.cur_left + cur.nxt_left == 0u) {
                pend = CLS_IDLE;
            }
        } else if (cls == CLS_ABSORB) {
            n_events += 1u;
            int ev;
            bin = interact_absorb(S, p, pending_container(pend), u_surf, w_z, ev);
            if (bin < 0) { n_events += 0x10000u; source = true; }   // re-emitted by the luminophore
        } else if (cls == CLS_FLAT || cls == CLS_CYL) {
            n_events += 1u;
            int ev;
            bin = interact_surface<EXT>(S, p, pending_kind(pend), pending_face(pend), pending_container(pend), pending_adj(pend),
                                        u_surf, ev);
            fly = bin < 0;
        } else if (cls == CLS_CELL) {
            fly = true;
        } else if (EXT && cls == CLS_OUT) {
            OutState q;
            q.x = abs_x(S, p); q.y = p.y; q.z = p.z + S.zc; q.dx = p.dx; q.dy = p.dy; q.dz = p.dz; q.aux = p.aux; q.steps = p.steps;
            const OutResult r = outside_step(S, q, u_surf);
            p.dx = r.dx; p.dy = r.dy; p.dz = r.dz; p.aux = r.aux; p.steps = r.steps;
            if (counts_as_event(r.event)) n_events += 1u + (unsigned)r.extra;
            if (r.entered) { p.y = r.y; locate_in_plate(S, p, r.x, r.z); }
            else { p.kc = 0; p.xr = r.x - axis_x(S, 0); p.y = r.y; p.z = r.z - S.zc; }
            bin = r.bin;
            fly = bin < 0;
        }
        // ---- source: a new photon or a re-emission --------------------------------------------------------
        if (source) {
            rng.batch = A.batches[batch].batch_id;
            const uint4 v = rng.next(A.rk);
            WavelengthDraw wd;
            if (fresh) bin = source_fresh<EXT, true>(S, A.batches[batch], A.spec_x, A.spec_cdf, A.spec_guide, rng, v, p, wd, nullptr, n_events);
            else source_emit(S, v, p, wd);
            source_finish(S, p, wd);
#ifdef SORT_SPLIT_SOURCE   // the advance of a (re-)emitted photon waits for the next deal: +1..4 % on the standard scene, -25 % with outside boxes
            if (bin < 0) pend = CLS_CELL;
#else
            fly = bin < 0;
#endif
        }
        // ---- common half of the next iteration ---------------------------------------------------------
        if (fly) {
            rng.batch = A.batches[batch].batch_id;
            const uint4 w = rng.next(A.rk);
            u_surf = u01(w.y); w_z = w.z;
            bin = advance<EXT, FLIGHT>(S, p, w.x, pend, n_events);
        }
        if (bin >= 0) { pend = CLS_FRESH; pbin = bin; }
    }
    if (hbatch >= 0) {
        unsigned long long* row = A.tallies + (size_t)hbatch * stride;
        for (int b = tid; b < stride; b += BLOCK) {
            const unsigned int v = hist[b];
            if (v) atomicAdd(row + b, (unsigned long long)v);
        }
    }
#ifdef SORT_TIMING
    if (lane == 0) {
        atomicAdd(&g_sort_timing[0], (unsigned long long)t_work); atomicAdd(&g_sort_timing[1], (unsigned long long)t_wait);
        atomicAdd(&g_sort_timing[2], (unsigned long long)t_sort); atomicAdd(&g_sort_timing[3], (unsigned long long)n_it);
        for (int c = 0; c < N_CLS; ++c) { atomicAdd(&g_sort_timing[8 + 2 * c], (unsigned long long)t_cls[c]); atomicAdd(&g_sort_timing[9 + 2 * c], (unsigned long long)n_cls[c]); }
        atomicAdd(&g_sort_timing[40], (unsigned long long)t_mixed); atomicAdd(&g_sort_timing[41], (unsigned long long)n_mixed);
    }
#endif
}

// K2p: warp-private photon pools, phase-major execution (lscpm_pool.cuh)
#define LSCPM_POOL_KERNEL_BODY
#include "lscpm_pool.cuh"

// K1 (debug/emit): the rays a batch generates, in plate-local coordinates
__global__ void emit_kernel(const __grid_constant__ DevScene S, const DevBatch* batch, const float* spec_x,
                            const float* spec_cdf, const unsigned short* spec_guide, unsigned long long photon_begin,
                            long long n, uint2 key, float* out /* [n][7]: origin, direction, wavelength */) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long g = photon_begin + (unsigned long long)i;
    Rng rng; rng.key = key; rng.p_lo = (uint32_t)g; rng.p_hi = (uint32_t)(g >> 32); rng.batch = batch->batch_id; rng.calls = 0;
    const uint4 v = rng.next();
    Photon p;
    WavelengthDraw wd;
    float origin[3];
    unsigned events;
    if (S.n_ext > 0) source_fresh<true, false>(S, *batch, spec_x, spec_cdf, spec_guide, rng, v, p, wd, origin, events);
    else source_fresh<false, false>(S, *batch, spec_x, spec_cdf, spec_guide, rng, v, p, wd, origin, events);
    source_finish(S, p, wd);
    float* o = out + 7 * i;
    o[0] = origin[0]; o[1] = origin[1]; o[2] = origin[2]; o[3] = p.dx; o[4] = p.dy; o[5] = p.dz; o[6] = p.wl;
}

// K5: deterministic per-ray probe — next interface and both outgoing directions
struct ProbeIn { float xr, y, z, dx, dy, dz; int kc, medium; };
struct ProbeOut {
    float t, n1, n2, R, cosi;
    float nx, ny, nz, rx, ry, rz, tx, ty, tz;
    int kind, face, ch, container, adj, found, ext;
};

__global__ void probe_kernel(const __grid_constant__ DevScene S, const ProbeIn* in, ProbeOut* out, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ProbeIn q = in[i];
    ProbeOut o;
    memset(&o, 0, sizeof(o));
    Photon p;
    p.xr = q.xr; p.y = q.y; p.z = q.z - S.zc; p.dx = q.dx; p.dy = q.dy; p.dz = q.dz; p.kc = q.kc; p.medium = q.medium;
    p.wl = 555.f; p.a_plate = p.a_outer = p.a_inner = 0.f; p.steps = 0; p.aux = 0;
    float nx, ny, nz;
    int ext_node = 0;      // probe bookkeeping: outside box hit (i+1)
    if (q.medium == MED_AIR && S.n_ext > 0) {
        float t; int node, face;
        const bool inside = outside_hit(S, abs_x(S, p), p.y, q.z, p.dx, p.dy, p.dz, -1, t, node, face);
        if (node < 0) { out[i] = o; return; }
        o.found = inside ? 2 : 1; o.t = t; o.kind = SURF_BOX; o.face = face; o.container = MED_AIR; o.adj = MED_PLATE; o.ch = node;
        ext_node = node;
        if (inside) {
            // the photon is inside box `node` (its zero-distance entry is dropped): pvtrace reports the box's exit face
            const DevExtBox& b = S.ext[node - 1];
            const float c[3] = {b.cx, b.cy, b.cz}, hh[3] = {b.hx, b.hy, b.hz}, oo[3] = {abs_x(S, p), p.y, q.z}, dd[3] = {p.dx, p.dy, p.dz};
            float tmax = LSCPM_INF; int fmax = 0;
            for (int ax = 0; ax < 3; ++ax) {
                if (dd[ax] == 0.f) continue;
                const float te = ((dd[ax] > 0.f ? hh[ax] : -hh[ax]) - (oo[ax] - c[ax])) / dd[ax];
                if (te < tmax) { tmax = te; fmax = 2 * ax + (dd[ax] > 0.f ? 1 : 0); }
            }
            o.t = tmax; o.face = face = fmax;
        }
        box_normal(face, nx, ny, nz);
    } else if (q.medium == MED_AIR) {
        const Entry e = plate_entry(S.hx, S.hy, S.hz, abs_x(S, p), p.y, q.z, p.dx, p.dy, p.dz);
        if (e.t < 0.f) { out[i] = o; return; }
        o.found = 1; o.t = e.t; o.kind = SURF_BOX; o.face = e.face; o.container = MED_AIR; o.adj = MED_PLATE; o.ch = 0;
        box_normal(e.face, nx, ny, nz);
    } else {
        // to the next real interface: the trace kernels' own geometry.  In the plastic of a scene without outside boxes
        // that is a plate flight without folding (every face ends it); otherwise compute_hit, marching over virtual cell
        // crossings.
        float travelled = 0.f;
        Hit h;
        if (S.n_ext == 0 && p.medium == MED_PLATE) {
            const float dx0 = p.dx, dy0 = p.dy, dz0 = p.dz;
            for (;;) {
                int cls, bounces;
                plate_flight<false>(S, p, LSCPM_INF, h, cls, bounces);
                if (cls != CLS_CELL) break;
                travelled += h.t;
            }
            // the flight moved the photon onto the interface: step back so that the common code below finds it at h.t
            p.xr = fmaf(-dx0, h.t, p.xr); p.y = fmaf(-dy0, h.t, p.y); p.z = fmaf(-dz0, h.t, p.z);
        } else {
            // coincident faces in scene-graph order (plate, tube, mixture), like the oracle's probe
            h = compute_hit<true, false>(S, p);
            while (h.kind == SURF_CELL) {
                travelled += h.t;
                p.y = fmaf(p.dy, h.t, p.y); p.z = fmaf(p.dz, h.t, p.z);
                p.kc += h.face; p.xr = h.face > 0 ? -0.5f * S.pitch : 0.5f * S.pitch;
                h = compute_hit<true, false>(S, p);
            }
        }
        const float xr = fmaf(p.dx, h.t, p.xr), yy = fmaf(p.dy, h.t, p.y), zz = fmaf(p.dz, h.t, p.z);
        o.found = 1; o.t = travelled + h.t; o.kind = h.kind; o.face = h.face; o.ch = p.kc; o.container = h.container; o.adj = h.adj;
        if (h.kind == SURF_BOX || h.kind == SURF_CAP) {
            box_normal(h.face, nx, ny, nz);
        } else {
            const float inv = rsqrtf(xr * xr + zz * zz);
            nx = xr * inv; ny = 0.f; nz = zz * inv;
        }
        (void)yy;
    }
    o.n1 = S.med[o.container].n; o.n2 = S.med[o.adj].n;
    if (ext_node > 0) o.n2 = S.ext[ext_node - 1].n;
    if (ext_node > 0 && o.found == 2) { o.n1 = S.ext[ext_node - 1].n; o.n2 = S.med[MED_AIR].n; }
    o.ext = ext_node;
    const Interface I = interface_eval(p.dx, p.dy, p.dz, nx, ny, nz, o.n1, o.n2);
    o.R = I.R; o.cosi = I.cosi; o.nx = I.nx; o.ny = I.ny; o.nz = I.nz;
    Photon r = p; reflect_dir(r, I);
    o.rx = r.dx; o.ry = r.dy; o.rz = r.dz;
    if (I.R < 1.f) { Photon t = p; refract_dir(t, I); o.tx = t.dx; o.ty = t.dy; o.tz = t.dz; }
    out[i] = o;
}

// follow with per-step recording (one thread per ray; used for photon paths and per-photon bins)
struct FollowIn { float ox, oy, oz, dx, dy, dz, wl; };
struct FollowStep { float x, y, z, dx, dy, dz, wl; int event, medium, ch, kind, face; };

__global__ void follow_kernel(const __grid_constant__ DevScene S, const FollowIn* in, long long n, uint32_t batch_id,
                              unsigned long long photon_begin, uint2 key, int max_steps, int* n_steps,
                              FollowStep* steps, int* final_bin) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FollowIn q = in[i];
    const unsigned long long g = photon_begin + (unsigned long long)i;
    Rng rng; rng.key = key; rng.p_lo = (uint32_t)g; rng.p_hi = (uint32_t)(g >> 32); rng.batch = batch_id; rng.calls = 0;
    Photon p;
    p.dx = q.dx; p.dy = q.dy; p.dz = q.dz; p.wl = q.wl;
    p.kc = 0; p.xr = 0.f; p.y = 0.f; p.z = 0.f; p.medium = MED_AIR; p.aux = 0;
    refresh_alphas(S, p);
    int k = 0;
    const bool ext = S.n_ext > 0;
    int bin = ext ? place_photon<true>(S, p, q.ox, q.oy, q.oz, false, 0.f, 0.f) : place_photon<false>(S, p, q.ox, q.oy, q.oz, false, 0.f, 0.f);
    while (bin < 0) {
        const uint4 w = rng.next();
        StepOut so;
        // recorders keep every reflection as a step of its own (FOLD = false)
        bin = ext ? photon_step<true, false, false>(S, p, w, so) : photon_step<false, true, false>(S, p, w, so);
        if (bin == STEP_EMIT) {
            const uint4 v = rng.next();
            WavelengthDraw wd;
            source_emit(S, v, p, wd);
            source_finish(S, p, wd);
            bin = -1;
        }
        if (so.event == EV_NONE) continue;
        if (k < max_steps) {
            FollowStep& s = steps[(size_t)i * max_steps + k];
            s.x = abs_x(S, p); s.y = p.y; s.z = p.z + S.zc; s.dx = p.dx; s.dy = p.dy; s.dz = p.dz; s.wl = p.wl;
            s.event = so.event; s.medium = so.medium; s.ch = so.ch; s.kind = so.kind; s.face = so.face;
            if (ext && so.event == EV_ABSORB && so.medium == MED_AIR && so.ch > 0) {
                // absorbed inside an opaque outside box (alpha ~1e10 m^-1: within 1e-10 m of its face).  The photon state
                // keeps the face point; the RECORD is placed 1e-7 m inside, so that the reference's
                // next_hit(scene, last_ray) (simulation_runner.py:45-53) starts inside the box it names - also for the side
                // PVs, whose entry face coincides with the plate's edge face (1e-10 m is below FP32 resolution here)
                s.x = fmaf(1e-7f, p.dx, s.x); s.y = fmaf(1e-7f, p.dy, s.y); s.z = fmaf(1e-7f, p.dz, s.z);
            }
        }
        ++k;
    }
    n_steps[i] = k;
    final_bin[i] = bin;
}

// =================================================================================================
// host side
// =================================================================================================
namespace {

thread_local std::string g_error;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const std::string& msg) { g_error = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    return fail(LSCPM_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CUDA_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return cuda_fail(e__, #expr); } while (0)

struct Vec3 { double x, y, z; };
inline Vec3 mul_rt(const double* pose, Vec3 v) {   // R^T v for a row-major 4x4 pose
    return {pose[0] * v.x + pose[4] * v.y + pose[8] * v.z, pose[1] * v.x + pose[5] * v.y + pose[9] * v.z,
            pose[2] * v.x + pose[6] * v.y + pose[10] * v.z};
}

// bin layout shared with lscpm_b200/tally.py and oracle/
struct BinLayout {
    int n_bins = 0;
    std::vector<int> exit_base, abs_base;
};
int n_components_of(const LscpmSceneDesc* d, int node) {
    const int m = d->node_material[node];
    return d->mat_comp_begin[m + 1] - d->mat_comp_begin[m];
}
BinLayout bin_layout(const LscpmSceneDesc* d) {
    BinLayout L;
    L.exit_base.assign(d->n_nodes, -1);
    L.abs_base.assign(d->n_nodes, -1);
    int next = 3;
    for (int i = 1; i < d->n_nodes; ++i) { L.exit_base[i] = next; next += 6; }
    for (int i = 0; i < d->n_nodes; ++i) { L.abs_base[i] = next; next += n_components_of(d, i); }
    L.n_bins = next;
    return L;
}
std::string bin_name(const LscpmSceneDesc* d, int bin) {
    static const char* box_faces[6] = {"-x", "+x", "-y", "+y", "-z", "+z"};
    static const char* cyl_faces[3] = {"side", "-cap", "+cap"};
    if (bin == 0) return "kill";
    if (bin == 1) return "exit/miss";
    if (bin == 2) return "exit/entry_reflect";
    const BinLayout L = bin_layout(d);
    if (bin < 0 || bin >= L.n_bins) return "";
    const int first_abs = L.abs_base[0];
    if (bin < first_abs) {
        const int node = 1 + (bin - 3) / 6, face = (bin - 3) % 6;
        std::string f;
        const int geom = d->node_geom[node];
        if (geom == LSCPM_GEOM_BOX) f = box_faces[face];
        else if (geom == LSCPM_GEOM_CYLINDER && face < 3) f = cyl_faces[face];
        else if (geom == LSCPM_GEOM_SPHERE && face == 0) f = "surface";
        else f = "unused" + std::to_string(face);
        return std::string("exit/") + d->node_name[node] + "/" + f;
    }
    for (int i = d->n_nodes - 1; i >= 0; --i) {
        if (bin >= L.abs_base[i] && bin < L.abs_base[i] + n_components_of(d, i)) {
            const int c = bin - L.abs_base[i];
            const int comp = d->mat_comp_begin[d->node_material[i]] + c;
            const char* prefix = d->comp_kind[comp] == LSCPM_COMP_REACTOR ? "react/" : "absorb/";
            return std::string(prefix) + d->node_name[i] + "/" + std::to_string(c);
        }
    }
    return "";
}

int validate_desc(const LscpmSceneDesc* d) {
    if (!d) return fail(LSCPM_EINVAL, "scene description is NULL");
    if (d->n_nodes < 1 || !d->node_geom || !d->node_parent || !d->node_material || !d->node_size || !d->node_pose ||
        !d->node_name)
        return fail(LSCPM_EINVAL, "scene description has no nodes or NULL node arrays");
    if (d->n_materials < 1 || !d->mat_refractive_index || !d->mat_comp_begin)
        return fail(LSCPM_EINVAL, "scene description has no materials");
    for (int i = 0; i < d->n_nodes; ++i) {
        if (d->node_material[i] < 0 || d->node_material[i] >= d->n_materials)
            return fail(LSCPM_EINVAL, "node material index out of range");
        if (d->node_parent[i] >= i) return fail(LSCPM_EINVAL, "nodes must be in pre-order (parent before child)");
        if (!d->node_name[i]) return fail(LSCPM_EINVAL, "node name is NULL");
    }
    if (d->mat_comp_begin[d->n_materials] != d->n_components) return fail(LSCPM_EINVAL, "component CSR does not match n_components");
    for (int c = 0; c < d->n_components; ++c) {
        if (d->comp_abs_table[c] >= d->n_tables || d->comp_ems_table[c] >= d->n_tables)
            return fail(LSCPM_EINVAL, "component table index out of range");
        if (d->comp_kind[c] == LSCPM_COMP_LUMINOPHORE && d->comp_ems_table[c] < 0)
            return fail(LSCPM_EINVAL, "luminophore without an emission table");
    }
    for (int t = 0; t < d->n_tables; ++t) {
        const int lo = d->table_begin[t], hi = d->table_begin[t + 1];
        if (hi - lo < 2) return fail(LSCPM_EINVAL, "spectrum table needs at least two knots");
        for (int i = lo + 1; i < hi; ++i)
            if (!(d->table_x[i] > d->table_x[i - 1])) return fail(LSCPM_EINVAL, "spectrum table abscissae must increase");
        for (int i = lo; i < hi; ++i)
            if (!(d->table_y[i] >= 0.0)) return fail(LSCPM_EINVAL, "spectrum table ordinates must be non-negative");
    }
    return LSCPM_OK;
}

// pvtrace Distribution: cdf = cumsum(y) / max(cdf) over the knots (knot spacing does not enter)
void cumulative_cdf(const double* x, const double* y, int n, std::vector<double>& cdf) {
    (void)x;
    cdf.assign(n, 0.0);
    double acc = 0.0, peak = 0.0;
    for (int i = 0; i < n; ++i) { acc += y[i]; cdf[i] = acc; peak = std::max(peak, acc); }
    if (peak > 0) for (int i = 0; i < n; ++i) cdf[i] /= peak;
}

bool same_material(const LscpmSceneDesc* d, int ma, int mb) {
    if (ma == mb) return true;
    if (d->mat_refractive_index[ma] != d->mat_refractive_index[mb]) return false;
    const int na = d->mat_comp_begin[ma + 1] - d->mat_comp_begin[ma], nb = d->mat_comp_begin[mb + 1] - d->mat_comp_begin[mb];
    if (na != nb) return false;
    for (int c = 0; c < na; ++c) {
        const int a = d->mat_comp_begin[ma] + c, b = d->mat_comp_begin[mb] + c;
        if (d->comp_kind[a] != d->comp_kind[b] || d->comp_coefficient[a] != d->comp_coefficient[b] ||
            d->comp_abs_table[a] != d->comp_abs_table[b] || d->comp_ems_table[a] != d->comp_ems_table[b] ||
            d->comp_quantum_yield[a] != d->comp_quantum_yield[b])
            return false;
    }
    return true;
}

}  // namespace

struct LscpmScene {
    int device = 0;
    DevScene dev{};
    // host copies needed to lower batches and to map results back to nodes / the world frame
    double plate_pose[16]{};
    double world_centre[3]{}, world_radius = 0;
    double hx = 0, hy = 0, hz = 0, cx0 = 0, pitch = 1, zc = 0, r_o = 0, r_i = 0;
    int n_ch = 0, has_inner = 0;
    int node_plate = 1;
    std::vector<int> node_outer, node_inner, node_ext;
    BinLayout layout;
    std::vector<std::string> bin_names;
    // device allocations
    float *d_tab_x = nullptr, *d_tab_y = nullptr, *d_tab_cdf = nullptr;
    unsigned short *d_tab_fwd = nullptr, *d_tab_inv = nullptr;
    int *d_abs_outer = nullptr, *d_abs_inner = nullptr, *d_exit_outer = nullptr, *d_exit_inner = nullptr;
    int axis_sign = 0;              // +1: the capillaries' local +z axis points along the plate's +y
    unsigned long long* d_counters = nullptr;   // ring of work counters, one per launch in flight
    std::atomic<unsigned> next_counter{0};
    // one event per ring slot, recorded behind the launch that used it: the launch that takes the slot over waits for it
    // (device-side, cudaStreamWaitEvent), so a counter is never zeroed under a kernel that still reads it
    cudaEvent_t counter_event[256] = {};
    std::mutex counter_mutex;
    int n_sm = 0, blocks_per_sm = 0;
    int trace_block = TRACE_BLOCK_BIG;   // threads per block of the lane kernel
    size_t trace_smem = 0;
    bool sorted = false;                 // block-sorted kernel or the lane-per-photon kernel
    bool sorted_forced = false;
    bool flight = false;                 // plate flights (plate holds a luminophore, no outside box shares one of its z faces)
    int sort_blocks_per_sm = 0;
    size_t sort_smem = 0;
    int pool = 0;                        // pool kernel configuration (PoolConfig), 0: lane / sorted kernel
    int pool_small_plan = 0;             // configuration for launches whose plan fits the small L1 (0: same as `pool`)
    // scratch of the host-buffer entry points (lscpm_trace_host, plan staging): one pinned host buffer, one device
    // buffer and one stream per scene, grown on demand and reused, so a call costs no allocation
    std::mutex scratch_mutex;
    unsigned char* h_scratch = nullptr;  // pinned
    unsigned char* d_scratch = nullptr;
    size_t h_scratch_bytes = 0, d_scratch_bytes = 0;
    cudaStream_t scratch_stream = nullptr;
};
constexpr int COUNTER_RING = 256;

struct LscpmPlan {
    LscpmScene* scene = nullptr;
    int n_batches = 0;
    unsigned char* d_blob = nullptr;     // ONE device allocation: lowered batches | spectrum knots | CDFs | guide tables
    bool owns_blob = true;
    DevBatch* d_batches = nullptr;
    float *d_spec_x = nullptr, *d_spec_cdf = nullptr;
    unsigned short* d_spec_guide = nullptr;
    uint64_t h2d_bytes = 0;
};

// host image of a plan: the four arrays laid out back to back (16-byte aligned sections) so that one copy uploads them
struct PlanLayout {
    size_t off_batches = 0, off_x = 0, off_cdf = 0, off_guide = 0, bytes = 0;
    size_t n_knots = 0, n_guide = 0;
};

namespace {

int lower_medium(const LscpmSceneDesc* d, int material, DevMedium& m) {
    m.n = (float)d->mat_refractive_index[material];
    const int c0 = d->mat_comp_begin[material];
    m.n_comp = d->mat_comp_begin[material + 1] - c0;
    if (m.n_comp > MAX_COMP) return fail(LSCPM_EUNSUPPORTED, "more than 4 components in one material");
    for (int c = 0; c < m.n_comp; ++c) {
        DevComponent& dc = m.comp[c];
        dc.kind = d->comp_kind[c0 + c];
        dc.abs_table = d->comp_abs_table[c0 + c];
        dc.ems_table = d->comp_ems_table[c0 + c];
        dc.coef = (float)d->comp_coefficient[c0 + c];
        dc.qy = (float)d->comp_quantum_yield[c0 + c];
        if (dc.kind != LSCPM_COMP_ABSORBER && dc.kind != LSCPM_COMP_LUMINOPHORE && dc.kind != LSCPM_COMP_REACTOR)
            return fail(LSCPM_EINVAL, "unknown component kind");
    }
    int n_tab = 0;
    m.single_tab = -1;
    double const_sum = 0;
    for (int c = 0; c < m.n_comp; ++c) {
        if (m.comp[c].abs_table >= 0) { ++n_tab; m.single_tab = c; }
        else const_sum += d->comp_coefficient[c0 + c];
    }
    if (n_tab != 1) m.single_tab = -1;
    m.const_sum = (float)const_sum;
    return LSCPM_OK;
}

// Recognise world > plate box > equally spaced y-parallel capillaries [> coaxial inner cylinder] and express it
// in the plate-local frame.  Everything is done in double; the kernels get the rounded floats.
int lower_scene(const LscpmSceneDesc* d, LscpmScene* s) {
    const double tol = 1e-9;
    if (d->n_nodes < 2) return fail(LSCPM_EUNSUPPORTED, "scene has no plate");
    if (d->node_geom[0] != LSCPM_GEOM_SPHERE) return fail(LSCPM_EUNSUPPORTED, "world node must be a sphere");
    if (n_components_of(d, 0) != 0) return fail(LSCPM_EUNSUPPORTED, "an absorbing world medium is not supported");
    int plate = -1;
    for (int i = 1; i < d->n_nodes; ++i)
        if (d->node_parent[i] == 0) {
            if (plate >= 0) return fail(LSCPM_EUNSUPPORTED, "more than one object directly inside the world");
            plate = i;
        }
    if (plate < 0 || d->node_geom[plate] != LSCPM_GEOM_BOX) return fail(LSCPM_EUNSUPPORTED, "the object inside the world must be a box (the plate)");
    s->node_plate = plate;
    s->world_radius = d->node_size[0];
    s->world_centre[0] = d->node_pose[3]; s->world_centre[1] = d->node_pose[7]; s->world_centre[2] = d->node_pose[11];
    const double* P = d->node_pose + 16 * plate;
    memcpy(s->plate_pose, P, sizeof(double) * 16);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double dot = 0;
            for (int k = 0; k < 3; ++k) dot += P[4 * k + r] * P[4 * k + c];
            if (std::fabs(dot - (r == c ? 1.0 : 0.0)) > 1e-9) return fail(LSCPM_EUNSUPPORTED, "plate pose is not rigid");
        }
    s->hx = 0.5 * d->node_size[3 * plate]; s->hy = 0.5 * d->node_size[3 * plate + 1]; s->hz = 0.5 * d->node_size[3 * plate + 2];

    struct Channel { double x, z; int outer, inner; };
    std::vector<Channel> ch;
    std::vector<Vec3> ext_centres, ext_halves;
    for (int i = 1; i < d->n_nodes; ++i) {
        if (i == plate || d->node_parent[i] != plate) continue;
        if (d->node_geom[i] == LSCPM_GEOM_BOX) {
            // a box in the plate's frame but outside its volume: the reference's photovoltaic absorbers
            // (miniplant/scene_creator.py:99-206).  Opaque (alpha * thickness >> 1), axis-aligned, touching allowed.
            const double* B = d->node_pose + 16 * i;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double rel = 0;
                    for (int k = 0; k < 3; ++k) rel += P[4 * k + r] * B[4 * k + c];
                    if (std::fabs(rel - (r == c ? 1.0 : 0.0)) > tol) return fail(LSCPM_EUNSUPPORTED, "outside boxes must be axis-aligned with the plate");
                }
            const Vec3 c = mul_rt(P, {B[3] - P[3], B[7] - P[7], B[11] - P[11]});
            const double h[3] = {0.5 * d->node_size[3 * i], 0.5 * d->node_size[3 * i + 1], 0.5 * d->node_size[3 * i + 2]};
            const double cc[3] = {c.x, c.y, c.z}, ph[3] = {s->hx, s->hy, s->hz};
            bool outside = false;
            for (int ax = 0; ax < 3; ++ax) if (std::fabs(cc[ax]) - h[ax] >= ph[ax] - tol) outside = true;
            if (!outside) return fail(LSCPM_EUNSUPPORTED, std::string("box '") + d->node_name[i] + "' overlaps the plate");
            const int m = d->node_material[i];
            const int c0 = d->mat_comp_begin[m], nc = d->mat_comp_begin[m + 1] - c0;
            const double thinnest = 2 * std::min(h[0], std::min(h[1], h[2]));
            if (nc != 1 || d->comp_kind[c0] != LSCPM_COMP_ABSORBER || d->comp_abs_table[c0] >= 0 || !(d->comp_coefficient[c0] * thinnest > 50.0))
                return fail(LSCPM_EUNSUPPORTED, std::string("box '") + d->node_name[i] + "' must be one opaque constant absorber");
            for (int j = i + 1; j < d->n_nodes; ++j)
                if (d->node_parent[j] == i) return fail(LSCPM_EUNSUPPORTED, "outside boxes cannot have children");
            if ((int)s->node_ext.size() >= MAX_EXT) return fail(LSCPM_EUNSUPPORTED, "more than 6 boxes outside the plate");
            s->node_ext.push_back(i);
            ext_centres.push_back(c); ext_halves.push_back({h[0], h[1], h[2]});
            continue;
        }
        if (d->node_geom[i] != LSCPM_GEOM_CYLINDER)
            return fail(LSCPM_EUNSUPPORTED, std::string("node '") + d->node_name[i] + "' is neither a capillary cylinder nor a box outside the plate");
        const double* C = d->node_pose + 16 * i;
        const Vec3 axis = mul_rt(P, {C[2], C[6], C[10]});                      // capillary z axis in plate coordinates
        const Vec3 centre = mul_rt(P, {C[3] - P[3], C[7] - P[7], C[11] - P[11]});
        if (std::fabs(std::fabs(axis.y) - 1.0) > tol || std::fabs(axis.x) > tol || std::fabs(axis.z) > tol)
            return fail(LSCPM_EUNSUPPORTED, "capillaries must run along the plate's y axis");
        if (std::fabs(centre.y) > tol || std::fabs(d->node_size[3 * i] - 2 * s->hy) > tol)
            return fail(LSCPM_EUNSUPPORTED, "capillaries must span the plate exactly (end caps on the +-y faces)");
        const int sign = axis.y > 0 ? 1 : -1;
        if (s->axis_sign != 0 && s->axis_sign != sign) return fail(LSCPM_EUNSUPPORTED, "capillaries must share the orientation of their axis");
        s->axis_sign = sign;
        Channel c{centre.x, centre.z, i, -1};
        for (int j = i + 1; j < d->n_nodes; ++j) {
            if (d->node_parent[j] != i) continue;
            if (c.inner >= 0 || d->node_geom[j] != LSCPM_GEOM_CYLINDER)
                return fail(LSCPM_EUNSUPPORTED, "a capillary may hold exactly one coaxial inner cylinder");
            const double* I = d->node_pose + 16 * j;
            for (int k = 0; k < 12; ++k)
                if (std::fabs(I[k] - C[k]) > tol) return fail(LSCPM_EUNSUPPORTED, "inner cylinder must be coaxial with its capillary");
            if (std::fabs(d->node_size[3 * j] - d->node_size[3 * i]) > tol) return fail(LSCPM_EUNSUPPORTED, "inner cylinder must have the capillary's length");
            for (int q = j + 1; q < d->n_nodes; ++q)
                if (d->node_parent[q] == j) return fail(LSCPM_EUNSUPPORTED, "nesting deeper than capillary > inner cylinder");
            c.inner = j;
        }
        ch.push_back(c);
    }
    for (int i = 1; i < d->n_nodes; ++i) {
        const int par = d->node_parent[i];
        if (i == plate || par == plate) continue;
        if (par <= 0 || d->node_parent[par] != plate) return fail(LSCPM_EUNSUPPORTED, "unexpected node outside the plate hierarchy");
        if (d->node_geom[par] != LSCPM_GEOM_CYLINDER) return fail(LSCPM_EUNSUPPORTED, "only capillaries may contain another node");
    }
    std::sort(ch.begin(), ch.end(), [](const Channel& a, const Channel& b) { return a.x < b.x; });
    s->n_ch = (int)ch.size();
    s->has_inner = 0;
    if (!ch.empty()) {
        const int o0 = ch[0].outer;
        s->r_o = d->node_size[3 * o0 + 1];
        s->has_inner = ch[0].inner >= 0;
        s->r_i = s->has_inner ? d->node_size[3 * ch[0].inner + 1] : 0.0;
        s->cx0 = ch[0].x; s->zc = ch[0].z;
        s->pitch = ch.size() > 1 ? (ch.back().x - ch[0].x) / (double)(ch.size() - 1) : 1.0;
        for (size_t k = 0; k < ch.size(); ++k) {
            const Channel& c = ch[k];
            if (std::fabs(c.x - (s->cx0 + k * s->pitch)) > tol || std::fabs(c.z - s->zc) > tol)
                return fail(LSCPM_EUNSUPPORTED, "capillaries must be equally spaced in x at a common height");
            if (std::fabs(d->node_size[3 * c.outer + 1] - s->r_o) > tol || !same_material(d, d->node_material[c.outer], d->node_material[o0]))
                return fail(LSCPM_EUNSUPPORTED, "capillaries must share radius and material");
            if ((c.inner >= 0) != (s->has_inner != 0)) return fail(LSCPM_EUNSUPPORTED, "either all or no capillaries have an inner cylinder");
            if (c.inner >= 0 && (std::fabs(d->node_size[3 * c.inner + 1] - s->r_i) > tol ||
                                 !same_material(d, d->node_material[c.inner], d->node_material[ch[0].inner])))
                return fail(LSCPM_EUNSUPPORTED, "inner cylinders must share radius and material");
            s->node_outer.push_back(c.outer);
            s->node_inner.push_back(c.inner);
        }
        if (s->has_inner && !(s->r_i < s->r_o)) return fail(LSCPM_EUNSUPPORTED, "inner radius must be smaller than the capillary radius");
        if (ch.size() > 1 && !(s->pitch > 2 * s->r_o)) return fail(LSCPM_EUNSUPPORTED, "capillaries overlap");
        if (!(std::fabs(s->zc) + s->r_o < s->hz) || !(std::fabs(ch[0].x) + s->r_o < s->hx) || !(std::fabs(ch.back().x) + s->r_o < s->hx))
            return fail(LSCPM_EUNSUPPORTED, "capillaries must lie strictly inside the plate");
    }

    DevScene& D = s->dev;
    memset(&D, 0, sizeof(D));
    D.hx = (float)s->hx; D.hy = (float)s->hy; D.hz = (float)s->hz;
    D.has_caps = s->n_ch > 0;
    D.has_inner = s->has_inner;
    if (s->n_ch > 0) {
        D.cx0 = (float)s->cx0; D.pitch = (float)s->pitch; D.zc = (float)s->zc; D.inv_pitch = (float)(1.0 / s->pitch);
        D.n_ch = s->n_ch;
        D.cell_lo_first = (float)(-s->hx - s->cx0);
        D.cell_hi_last = (float)(s->hx - (s->cx0 + (s->n_ch - 1) * s->pitch));
        D.r_o = (float)s->r_o; D.r_o2 = (float)(s->r_o * s->r_o);
        D.r_i = (float)s->r_i; D.r_i2 = s->has_inner ? (float)(s->r_i * s->r_i) : -1.f;
    } else {
        // a bare plate is one cell without a capillary: negative squared radii never produce a root
        D.cx0 = 0.f; D.pitch = (float)(4 * s->hx); D.zc = 0.f; D.inv_pitch = (float)(1.0 / (4 * s->hx));
        D.n_ch = 1;
        D.cell_lo_first = (float)(-s->hx); D.cell_hi_last = (float)s->hx;
        D.r_o = D.r_i = 0.f; D.r_o2 = D.r_i2 = -1.f;
    }
    D.med[MED_AIR].n = (float)d->mat_refractive_index[d->node_material[0]];
    D.med[MED_AIR].n_comp = 0;
    int rc = lower_medium(d, d->node_material[plate], D.med[MED_PLATE]);
    if (rc) return rc;
    if (!ch.empty()) {
        if ((rc = lower_medium(d, d->node_material[ch[0].outer], D.med[MED_OUTER]))) return rc;
        if (s->has_inner && (rc = lower_medium(d, d->node_material[ch[0].inner], D.med[MED_INNER]))) return rc;
    }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) D.w2l[3 * r + c] = (float)P[4 * c + r];   // R^T
    s->layout = bin_layout(d);
    D.n_bins = s->layout.n_bins;
    D.exit_base_plate = s->layout.exit_base[plate];
    D.abs_base_plate = s->layout.abs_base[plate];
    D.med[MED_OUT] = D.med[MED_AIR];
    D.n_ext = (int)s->node_ext.size();
    for (int e = 0; e < D.n_ext; ++e) {
        DevExtBox& b = D.ext[e];
        const int node = s->node_ext[e];
        b.cx = (float)ext_centres[e].x; b.cy = (float)ext_centres[e].y; b.cz = (float)ext_centres[e].z;
        b.hx = (float)ext_halves[e].x; b.hy = (float)ext_halves[e].y; b.hz = (float)ext_halves[e].z;
        b.n = (float)d->mat_refractive_index[d->node_material[node]];
        b.abs_bin = s->layout.abs_base[node];
        b.exit_base = s->layout.exit_base[node];
    }
    for (int f = 0; f < 6; ++f) {   // boxes that share plate face f: their opposite face lies in its plane
        D.ext_touch[f] = 0u;
        const double plate_half[3] = {(double)D.hx, (double)D.hy, (double)D.hz};
        const int ax = f / 2;
        for (int e = 0; e < D.n_ext; ++e) {
            const double c[3] = {ext_centres[e].x, ext_centres[e].y, ext_centres[e].z}, h[3] = {ext_halves[e].x, ext_halves[e].y, ext_halves[e].z};
            const double gap = (f & 1) ? (c[ax] - h[ax]) - plate_half[ax] : -plate_half[ax] - (c[ax] + h[ax]);
            if (std::fabs(gap) <= (double)EXT_TOL) D.ext_touch[f] |= 1u << e;
        }
    }
    s->bin_names.resize(D.n_bins);
    for (int b = 0; b < D.n_bins; ++b) s->bin_names[b] = bin_name(d, b);
    return LSCPM_OK;
}

template <typename T>
int upload(const std::vector<T>& host, T** dev) {
    *dev = nullptr;
    const size_t bytes = std::max<size_t>(host.size(), 1) * sizeof(T);
    CUDA_TRY(cudaMalloc((void**)dev, bytes));
    if (!host.empty()) CUDA_TRY(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return LSCPM_OK;
}

// world -> plate-local helpers (double)
inline void world_point_to_plate(const double* P, const double* w, double* out) {
    const Vec3 v = mul_rt(P, {w[0] - P[3], w[1] - P[7], w[2] - P[11]});
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}
inline void world_vec_to_plate(const double* P, const double* w, double* out) {
    const Vec3 v = mul_rt(P, {w[0], w[1], w[2]});
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}
inline void plate_vec_to_world(const double* P, const double* l, double* out) {
    for (int r = 0; r < 3; ++r) out[r] = P[4 * r] * l[0] + P[4 * r + 1] * l[1] + P[4 * r + 2] * l[2];
}
inline void plate_point_to_world(const double* P, const double* l, double* out) {
    plate_vec_to_world(P, l, out);
    out[0] += P[3]; out[1] += P[7]; out[2] += P[11];
}

int lower_batch(const LscpmScene* s, const LscpmBatchDesc* b, const std::vector<int>& spec_begin, DevBatch& o) {
    memset(&o, 0, sizeof(o));
    if (b->light_kind != LSCPM_LIGHT_DIRECT && b->light_kind != LSCPM_LIGHT_DIFFUSE) return fail(LSCPM_EINVAL, "unknown light kind");
    o.light_kind = b->light_kind;
    o.batch_id = b->batch_id;
    o.wavelength = (float)b->wavelength_nm;
    if (b->spectrum >= 0) {
        if (b->spectrum + 1 >= (int)spec_begin.size()) return fail(LSCPM_EINVAL, "batch spectrum index out of range");
        o.spec_begin = spec_begin[b->spectrum];
        o.spec_n = spec_begin[b->spectrum + 1] - spec_begin[b->spectrum];
        o.spec_guide = b->spectrum * (SPEC_BUCKETS + 1);
    } else {
        o.spec_begin = -1; o.spec_n = 0;
        if (!(b->wavelength_nm > 0)) return fail(LSCPM_EINVAL, "constant wavelength must be positive");
    }
    if (!(b->mask_half[0] >= 0) || !(b->mask_half[1] >= 0) || !(b->back_off >= 0)) return fail(LSCPM_EINVAL, "negative mask extent or back-off");
    o.mask_half[0] = (float)b->mask_half[0]; o.mask_half[1] = (float)b->mask_half[1];
    const double* P = s->plate_pose;
    double A[12];
    for (int c = 0; c < 4; ++c) {
        const double col[3] = {b->mask_pose[c], b->mask_pose[4 + c], b->mask_pose[8 + c]};
        double out[3];
        if (c < 3) world_vec_to_plate(P, col, out); else world_point_to_plate(P, col, out);
        for (int r = 0; r < 3; ++r) A[4 * r + c] = out[r];
    }
    for (int k = 0; k < 12; ++k) o.A[k] = (float)A[k];
    if (b->light_kind == LSCPM_LIGHT_DIRECT) {
        const double n2 = b->direction[0] * b->direction[0] + b->direction[1] * b->direction[1] + b->direction[2] * b->direction[2];
        if (std::fabs(n2 - 1.0) > 1e-9) return fail(LSCPM_EINVAL, "beam direction must be a unit vector");
        double dl[3];
        world_vec_to_plate(P, b->direction, dl);
        for (int k = 0; k < 3; ++k) o.dir[k] = (float)dl[k];
    }
    o.back_off = (float)b->back_off;
    const double tilt = b->tilt_deg * M_PI / 180.0;
    o.cos_tilt = (float)std::cos(tilt); o.sin_tilt = (float)std::sin(tilt);
    // anchors on the top face?  (standard LSC-PM lights: reference utils.py:87-99 puts them exactly there)
    const double tol = 1e-9;
    const bool flat = std::fabs(A[8]) < tol && std::fabs(A[9]) < tol && std::fabs(A[11] - s->hz) < tol;
    const double ex = std::fabs(A[0]) * b->mask_half[0] + std::fabs(A[1]) * b->mask_half[1] + std::fabs(A[3]);
    const double ey = std::fabs(A[4]) * b->mask_half[0] + std::fabs(A[5]) * b->mask_half[1] + std::fabs(A[7]);
    o.anchor_on_top = flat && ex <= s->hx + tol && ey <= s->hy + tol && b->back_off > 1e-6;
    return LSCPM_OK;
}

typedef void (*TraceKernel)(const DevScene, const TraceArgs);
TraceKernel lane_kernel(const LscpmScene* s) {
    if (s->trace_block == TRACE_BLOCK_BIG) {
        if (s->dev.n_ext > 0) return s->flight ? trace_kernel<true, true, TRACE_BLOCK_BIG> : trace_kernel<true, false, TRACE_BLOCK_BIG>;
        return s->flight ? trace_kernel<false, true, TRACE_BLOCK_BIG> : trace_kernel<false, false, TRACE_BLOCK_BIG>;
    }
    if (s->dev.n_ext > 0) return s->flight ? trace_kernel<true, true, TRACE_BLOCK_SMALL> : trace_kernel<true, false, TRACE_BLOCK_SMALL>;
    return s->flight ? trace_kernel<false, true, TRACE_BLOCK_SMALL> : trace_kernel<false, false, TRACE_BLOCK_SMALL>;
}
TraceKernel sorted_kernel(const LscpmScene* s) {
    if (s->dev.n_ext > 0) return s->flight ? trace_sorted_kernel<true, true, SORT_THREADS> : trace_sorted_kernel<true, false, SORT_THREADS>;
    return s->flight ? trace_sorted_kernel<false, true, SORT_THREADS> : trace_sorted_kernel<false, false, SORT_THREADS>;
}

// pool kernel configurations: records per warp decide the lanes per phase execution (120: 31.8 of 32, 96: 31, 80: 29,
// 64: 26) and are limited by shared memory; what shared memory takes, the L1 loses (up to 196 KB leaves 60 KB of L1, more
// leaves 28 KB).  Measured (one B200, ms per job; profiles/r2_pool_records_ab.txt; end state of round 2):
//                                              128 rec.   120      96
//   config 2 (64 batches, one spectrum)          7.06     7.06    7.15     (K2: 10.0)
//   config 4 (16 096 batches, 16 096 spectra)   189.4    189.4   183.7
//   config 5 (dense grid: 702 counters per row)    (do not fit)   7.87     (K2: 11.7)
// so: 96 records by default (fewer for scenes whose per-warp tally rows are very wide), 120 for launches whose batch
// descriptors and spectra are small enough to live in the 28 KB L1 beside the scene tables.  With the five-phase version
// of the kernel: 80 records 8.05 ms against 96 at 7.67 on config 2, and a 28-warp x 80-record configuration (72 registers)
// 7.94 ms: more warps do not buy back the lanes fewer records lose.
constexpr int POOL_WARPS = 24, POOL_WARPS_WIDE = 32;
constexpr size_t POOL_SMALL_PLAN_BYTES = 16 * 1024;
enum PoolConfig { POOL_OFF = 0, POOL_96 = 1, POOL_WIDE_64 = 2, POOL_80 = 3, POOL_64 = 4, POOL_120 = 7 };
int pool_warps(int cfg) { return cfg == POOL_WIDE_64 ? POOL_WARPS_WIDE : POOL_WARPS; }
int pool_slots(int cfg) { return (cfg == POOL_120 ? 120 : (cfg == POOL_96 ? 96 : (cfg == POOL_80 ? 80 : 64))); }
TraceKernel pool_kernel(const LscpmScene* s, int cfg) {
    if (s->dev.n_ext > 0) {   // scenes with outside boxes: the air phase in the plate flights' list slot, or in a fifth list
        switch (cfg) {
            case POOL_120: return s->flight ? trace_pool_kernel<true, true, POOL_WARPS, 120> : trace_pool_kernel<true, false, POOL_WARPS, 120>;
            case POOL_96: return s->flight ? trace_pool_kernel<true, true, POOL_WARPS, 96> : trace_pool_kernel<true, false, POOL_WARPS, 96>;
            case POOL_80: return s->flight ? trace_pool_kernel<true, true, POOL_WARPS, 80> : trace_pool_kernel<true, false, POOL_WARPS, 80>;
            default: return s->flight ? trace_pool_kernel<true, true, POOL_WARPS, 64> : trace_pool_kernel<true, false, POOL_WARPS, 64>;
        }
    }
    switch (cfg) {
        case POOL_120: return s->flight ? trace_pool_kernel<false, true, POOL_WARPS, 120> : trace_pool_kernel<false, false, POOL_WARPS, 120>;
        case POOL_96: return s->flight ? trace_pool_kernel<false, true, POOL_WARPS, 96> : trace_pool_kernel<false, false, POOL_WARPS, 96>;
        case POOL_80: return s->flight ? trace_pool_kernel<false, true, POOL_WARPS, 80> : trace_pool_kernel<false, false, POOL_WARPS, 80>;
        case POOL_64: return s->flight ? trace_pool_kernel<false, true, POOL_WARPS, 64> : trace_pool_kernel<false, false, POOL_WARPS, 64>;
        default: return s->flight ? trace_pool_kernel<false, true, POOL_WARPS_WIDE, 64> : trace_pool_kernel<false, false, POOL_WARPS_WIDE, 64>;
    }
}
size_t pool_smem(const LscpmScene* s, int cfg) {
    return pool_smem_bytes(pool_warps(cfg), pool_slots(cfg), s->dev.n_bins + LSCPM_N_COUNTERS, pool_lists(s->dev.n_ext > 0, s->flight));
}

int pick_counter(LscpmScene* s, cudaStream_t stream, unsigned long long** out, unsigned* slot_out) {
    const unsigned slot = s->next_counter.fetch_add(1) % COUNTER_RING;
    *out = s->d_counters + slot;
    *slot_out = slot;
    {
        std::lock_guard<std::mutex> lock(s->counter_mutex);
        if (s->counter_event[slot]) CUDA_TRY(cudaStreamWaitEvent(stream, s->counter_event[slot], 0));
        else CUDA_TRY(cudaEventCreateWithFlags(&s->counter_event[slot], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaMemsetAsync(*out, 0, sizeof(unsigned long long), stream));
    return LSCPM_OK;
}

int release_counter(LscpmScene* s, cudaStream_t stream, unsigned slot) {
    std::lock_guard<std::mutex> lock(s->counter_mutex);
    CUDA_TRY(cudaEventRecord(s->counter_event[slot], stream));
    return LSCPM_OK;
}

}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

int lscpm_abi_version(void) { return LSCPM_ABI_VERSION; }
const char* lscpm_last_error(void) { return g_error.c_str(); }
uint64_t lscpm_launch_count(void) { return g_launches.load(); }

void lscpm_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    const uint4 r = philox4x32_10(make_uint4(counter[0], counter[1], counter[2], counter[3]), make_uint2(key[0], key[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

int lscpm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lscpm_desc_n_bins(const LscpmSceneDesc* desc) {
    const int rc = validate_desc(desc);
    if (rc) return rc;
    return bin_layout(desc).n_bins;
}

int lscpm_desc_bin_name(const LscpmSceneDesc* desc, int bin, char* buf, int buf_len) {
    const int rc = validate_desc(desc);
    if (rc) return rc;
    const std::string name = bin_name(desc, bin);
    if (name.empty()) return fail(LSCPM_EINVAL, "bin index out of range");
    if (buf && buf_len > 0) { strncpy(buf, name.c_str(), (size_t)buf_len - 1); buf[buf_len - 1] = 0; }
    return (int)name.size();
}

int lscpm_scene_create(const LscpmSceneDesc* desc, int device, LscpmScene** out) {
    if (!out) return fail(LSCPM_EINVAL, "out pointer is NULL");
    *out = nullptr;
    int rc = validate_desc(desc);
    if (rc) return rc;
    LscpmScene* s = new (std::nothrow) LscpmScene();
    if (!s) return fail(LSCPM_ENOMEM, "out of host memory");
    if ((rc = lower_scene(desc, s))) { delete s; return rc; }
    auto bail = [&](int code) { lscpm_scene_destroy(s); return code; };
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        delete s;
        return fail(LSCPM_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n_dev) { delete s; return fail(LSCPM_EINVAL, "device index out of range"); }
    s->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { delete s; return cuda_fail(e, "cudaSetDevice"); }

    std::vector<float> tx, ty, tc;
    std::vector<unsigned short> fwd, inv;
    int n_dev_tables = 0;
    // Stores one knot table (resampled onto its regular grid when the knots sit on one) and returns its index.
    auto add_table = [&](const std::vector<double>& x, const std::vector<double>& y, int* index) -> int {
        const int n = (int)x.size();
        if (n_dev_tables >= MAX_TABLES) return fail(LSCPM_EUNSUPPORTED, "too many spectrum tables in one scene");
        std::vector<double> cdf;
        cumulative_cdf(x.data(), y.data(), n, cdf);
        double wmin = x[n - 1] - x[0];
        for (int i = 1; i < n; ++i) wmin = std::min(wmin, x[i] - x[i - 1]);
        const double range = x[n - 1] - x[0];
        // regular grid?  every knot within 1e-6 of a multiple of the smallest spacing
        bool regular = range / wmin < 60000.0;
        for (int i = 0; regular && i < n; ++i) {
            const double q = (x[i] - x[0]) / wmin;
            if (std::fabs(q - std::round(q)) > 1e-6) regular = false;
        }
        std::vector<double> gx, gy, gc;
        if (regular) {
            const int m = (int)std::llround(range / wmin) + 1;
            int j = 0;
            for (int i = 0; i < m; ++i) {
                const double xi = i == m - 1 ? x[n - 1] : x[0] + i * wmin;
                while (j + 2 < n && x[j + 1] <= xi) ++j;
                const double f = (xi - x[j]) / (x[j + 1] - x[j]);
                gx.push_back(xi); gy.push_back(y[j] + f * (y[j + 1] - y[j])); gc.push_back(cdf[j] + f * (cdf[j + 1] - cdf[j]));
            }
        } else { gx = x; gy = y; gc = cdf; }
        const int m = (int)gx.size();
        if (m > 65000) return fail(LSCPM_EUNSUPPORTED, "spectrum table with more than 65000 knots");
        DevTable& T = s->dev.tab[n_dev_tables];
        T.begin = (int)tx.size(); T.n = m; T.uniform = regular ? 1 : 0;
        for (int i = 0; i < m; ++i) { tx.push_back((float)gx[i]); ty.push_back((float)gy[i]); tc.push_back((float)gc[i]); }
        T.x0 = (float)gx[0]; T.fwd_begin = (int)fwd.size();
        if (regular) { T.inv_w = (float)(1.0 / wmin); T.fwd_n = 0; }
        else {
            int nb = std::max(1, std::min((int)std::ceil(range / wmin), 16384));
            const double w = range / nb;
            T.inv_w = (float)(1.0 / w); T.fwd_n = nb;
            int j = 0;
            for (int bkt = 0; bkt < nb; ++bkt) {
                const double start = gx[0] + bkt * w;
                while (j + 2 < m && gx[j + 1] <= start) ++j;
                fwd.push_back((unsigned short)j);
            }
        }
        // inverse guide on the CDF: idx[b] = last knot (<= m-2) with cdf <= b / G
        T.inv_begin = (int)inv.size();
        int j = 0;
        for (int bkt = 0; bkt <= INV_CDF_BUCKETS; ++bkt) {
            const double level = (double)bkt / INV_CDF_BUCKETS;
            while (j + 2 < m && gc[j + 1] <= level) ++j;
            inv.push_back((unsigned short)j);
        }
        *index = n_dev_tables++;
        return LSCPM_OK;
    };
    auto knots_of = [&](int t, std::vector<double>& x, std::vector<double>& y) {
        const int lo = desc->table_begin[t], n = desc->table_begin[t + 1] - lo;
        x.assign(desc->table_x + lo, desc->table_x + lo + n);
        y.assign(desc->table_y + lo, desc->table_y + lo + n);
    };
    auto interp_clamped = [](const std::vector<double>& x, const std::vector<double>& y, double v) {
        if (v <= x.front()) return y.front();
        if (v >= x.back()) return y.back();
        const size_t j = std::upper_bound(x.begin(), x.end(), v) - x.begin() - 1;
        return y[j] + (v - x[j]) / (x[j + 1] - x[j]) * (y[j + 1] - y[j]);
    };
    for (int t = 0; t < desc->n_tables; ++t) {   // scene tables keep their indices
        std::vector<double> x, y; int idx;
        knots_of(t, x, y);
        if ((rc = add_table(x, y, &idx))) return bail(rc);
    }
    // one total-attenuation table per medium: constants folded in, knots = union of the components' knots
    for (int med = MED_PLATE; med <= MED_INNER; ++med) {
        DevMedium& m = s->dev.med[med];
        m.alpha_table = -1;
        std::vector<double> ux;
        for (int c = 0; c < m.n_comp; ++c)
            if (m.comp[c].abs_table >= 0) {
                std::vector<double> x, y; knots_of(m.comp[c].abs_table, x, y);
                ux.insert(ux.end(), x.begin(), x.end());
            }
        if (ux.empty()) continue;
        std::sort(ux.begin(), ux.end());
        ux.erase(std::unique(ux.begin(), ux.end()), ux.end());
        std::vector<double> uy(ux.size(), (double)0);
        for (size_t i = 0; i < ux.size(); ++i) {
            double total = 0;
            const int mat_c0 = med == MED_PLATE ? desc->mat_comp_begin[desc->node_material[s->node_plate]]
                             : med == MED_OUTER ? desc->mat_comp_begin[desc->node_material[s->node_outer[0]]]
                                                : desc->mat_comp_begin[desc->node_material[s->node_inner[0]]];
            for (int c = 0; c < m.n_comp; ++c) {
                if (m.comp[c].abs_table < 0) total += desc->comp_coefficient[mat_c0 + c];
                else { std::vector<double> x, y; knots_of(m.comp[c].abs_table, x, y); total += interp_clamped(x, y, ux[i]); }
            }
            uy[i] = total;
        }
        if ((rc = add_table(ux, uy, &m.alpha_table))) return bail(rc);
    }
    std::vector<int> abs_outer, abs_inner, exit_outer, exit_inner;
    for (int k = 0; k < s->n_ch; ++k) {
        abs_outer.push_back(s->layout.abs_base[s->node_outer[k]]);
        abs_inner.push_back(s->node_inner[k] >= 0 ? s->layout.abs_base[s->node_inner[k]] : 0);
        exit_outer.push_back(s->layout.exit_base[s->node_outer[k]]);
        exit_inner.push_back(s->node_inner[k] >= 0 ? s->layout.exit_base[s->node_inner[k]] : 0);
    }
    // cylinder face slots: 0 side, 1 -cap (local -z), 2 +cap (local +z); which of them lies in the plate's +y face
    s->dev.cap_slot_hi = s->axis_sign >= 0 ? 2 : 1;
    s->dev.cap_slot_lo = s->axis_sign >= 0 ? 1 : 2;
    if ((rc = upload(tx, &s->d_tab_x)) || (rc = upload(ty, &s->d_tab_y)) || (rc = upload(tc, &s->d_tab_cdf)) ||
        (rc = upload(fwd, &s->d_tab_fwd)) || (rc = upload(inv, &s->d_tab_inv)) ||
        (rc = upload(abs_outer, &s->d_abs_outer)) || (rc = upload(abs_inner, &s->d_abs_inner)) ||
        (rc = upload(exit_outer, &s->d_exit_outer)) || (rc = upload(exit_inner, &s->d_exit_inner)))
        return bail(rc);
    if ((e = cudaMalloc((void**)&s->d_counters, sizeof(unsigned long long) * COUNTER_RING)) != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc counters"));
    s->dev.tab_x = s->d_tab_x; s->dev.tab_y = s->d_tab_y; s->dev.tab_cdf = s->d_tab_cdf;
    s->dev.tab_fwd = s->d_tab_fwd; s->dev.tab_inv = s->d_tab_inv;
    s->dev.abs_base_outer = s->d_abs_outer; s->dev.abs_base_inner = s->d_abs_inner;
    s->dev.exit_base_outer = s->d_exit_outer; s->dev.exit_base_inner = s->d_exit_inner;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(cuda_fail(e, "cudaGetDeviceProperties"));
    s->n_sm = prop.multiProcessorCount;
    // plate flights pay where light is trapped in the plastic, i.e. when the plate's material re-emits; a plate without a
    // luminophore (the reference's include_dye=False) is crossed once and is faster with the plain cell march.
    // LSCPM_PLATE_FLIGHTS=0|1 overrides (diagnostics; tallies agree statistically, not bit for bit).
    // A flight folds the reflections at the z faces with the test against the world's medium: an outside box that shares
    // one of those faces (none of the reference's does: the bottom cells hang 17 mm below the plate) rules flights out.
    const bool flights_valid = s->dev.ext_touch[4] == 0u && s->dev.ext_touch[5] == 0u;
    s->flight = false;
    if (flights_valid)
        for (int c = 0; c < s->dev.med[MED_PLATE].n_comp; ++c)
            if (s->dev.med[MED_PLATE].comp[c].kind == COMP_LUMINOPHORE) s->flight = true;
    if (const char* f = getenv("LSCPM_PLATE_FLIGHTS")) s->flight = flights_valid && f[0] == '1';
#ifdef LSCPM_DIAG
    s->dev.diag_emit = getenv("LSCPM_DIAG_EMIT") ? atoi(getenv("LSCPM_DIAG_EMIT")) : 0;
#endif
    const size_t row_bytes = (size_t)(s->dev.n_bins + LSCPM_N_COUNTERS) * sizeof(unsigned int);
    s->trace_block = (TRACE_BLOCK_BIG / 32) * row_bytes <= TRACE_SMEM_BIG_LIMIT ? TRACE_BLOCK_BIG : TRACE_BLOCK_SMALL;
    if (const char* b = getenv("LSCPM_TRACE_BLOCK")) s->trace_block = atoi(b) == TRACE_BLOCK_SMALL ? TRACE_BLOCK_SMALL : TRACE_BLOCK_BIG;
    s->trace_smem = (size_t)(s->trace_block / 32) * row_bytes;
    if (s->trace_smem > 200 * 1024) return bail(fail(LSCPM_EUNSUPPORTED, "too many tally bins for the per-warp shared rows"));
    // the opt-in limit is a property of the kernel, shared by every scene handle of the process: raise it to the most any
    // scene may ask of this instantiation, never to one scene's own need (a later, smaller scene would lower it again)
    if (s->trace_smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(lane_kernel(s), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  s->trace_block == TRACE_BLOCK_BIG ? (int)TRACE_SMEM_BIG_LIMIT : 200 * 1024)) != cudaSuccess)
        return bail(cuda_fail(e, "cudaFuncSetAttribute"));
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->blocks_per_sm, lane_kernel(s),
                                                           s->trace_block, s->trace_smem)) != cudaSuccess)
        return bail(cuda_fail(e, "occupancy query (is the library built for this GPU's architecture?)"));
    if (s->blocks_per_sm < 1) return bail(fail(LSCPM_ECUDA, "trace kernel does not fit on an SM"));
    // The block-sorted kernel wins where photons do very different amounts of work per iteration (scenes with outside
    // boxes: +40..60 %); on the standard scene the two are within a few per cent and the lane kernel has no barriers.
    // LSCPM_TRACE_KERNEL=lanes|sorted forces one (the tallies are bit-identical either way).
    const char* which = getenv("LSCPM_TRACE_KERNEL");
    s->sorted = s->dev.n_ext > 0;
    if (which && strcmp(which, "lanes") == 0) s->sorted = false;
    if (which && strcmp(which, "sorted") == 0) s->sorted = true;
    s->sorted_forced = which && strcmp(which, "sorted") == 0;
    s->sort_smem = (size_t)(s->dev.n_bins + LSCPM_N_COUNTERS) * sizeof(unsigned int);
    if (s->sort_smem > 24 * 1024) s->sorted = false;    // static exchange buffer + row must stay under 48 KB
    if (s->sorted) {
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                 &s->sort_blocks_per_sm, sorted_kernel(s),
                 SORT_THREADS, s->sort_smem)) != cudaSuccess)
            return bail(cuda_fail(e, "occupancy query (sorted kernel)"));
        if (s->sort_blocks_per_sm < 1) return bail(fail(LSCPM_ECUDA, "sorted trace kernel does not fit on an SM"));
    }
    // The pool kernel (warp-private photon pools, lscpm_pool.cuh) is the default for scenes whose plate re-emits - that is
    // where photons live long enough for the phase lists to stay full - and for every scene with outside boxes.
    // LSCPM_TRACE_KERNEL=pool|pool96|pool80|pool64 forces it (bit-identical tallies in every configuration).
    s->pool = POOL_OFF;
    s->pool_small_plan = POOL_OFF;
    {
        const bool ext = s->dev.n_ext > 0;
        const size_t optin = (size_t)prop.sharedMemPerBlockOptin;
        auto fits = [&](int cfg) { return pool_smem(s, cfg) <= optin; };
        const bool forced = which && strncmp(which, "pool", 4) == 0;
        bool pinned = false;        // a configuration named in the environment is used for every launch
        if (s->flight || forced || ext) s->pool = POOL_96;
        if (which && (strcmp(which, "lanes") == 0 || strcmp(which, "sorted") == 0)) s->pool = POOL_OFF;
        if (which && strcmp(which, "pool120") == 0) { s->pool = POOL_120; pinned = true; }
        if (which && strcmp(which, "pool96") == 0) { s->pool = POOL_96; pinned = true; }
        if (which && strcmp(which, "pool80") == 0) { s->pool = POOL_80; pinned = true; }
        if (which && strcmp(which, "pool64") == 0) { s->pool = ext ? POOL_64 : POOL_WIDE_64; pinned = true; }
        while (s->pool && !fits(s->pool))      // wide tally rows: fewer records, finally the lane / sorted kernel
            s->pool = (s->pool == POOL_120) ? POOL_96 : (s->pool == POOL_96 ? POOL_80 : (s->pool == POOL_80 ? POOL_64 : POOL_OFF));
        if (s->pool == POOL_96 && !pinned && fits(POOL_120)) s->pool_small_plan = POOL_120;
        for (int cfg : {s->pool, s->pool_small_plan}) {
            if (!cfg) continue;
            if ((e = cudaFuncSetAttribute(pool_kernel(s, cfg), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin)) != cudaSuccess)
                return bail(cuda_fail(e, "cudaFuncSetAttribute (pool kernel)"));
            int fit = 0;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, pool_kernel(s, cfg), pool_warps(cfg) * 32, pool_smem(s, cfg))) != cudaSuccess)
                return bail(cuda_fail(e, "occupancy query (pool kernel)"));
            if (fit < 1) { if (cfg == s->pool) s->pool = POOL_OFF; s->pool_small_plan = POOL_OFF; }
        }
        if (s->pool) s->sorted = false;
    }
    *out = s;
    return LSCPM_OK;
}

int lscpm_scene_destroy(LscpmScene* s) {
    if (!s) return LSCPM_OK;
    cudaFree(s->d_tab_x); cudaFree(s->d_tab_y); cudaFree(s->d_tab_cdf); cudaFree(s->d_tab_fwd); cudaFree(s->d_tab_inv);
    cudaFree(s->d_abs_outer); cudaFree(s->d_abs_inner); cudaFree(s->d_exit_outer); cudaFree(s->d_exit_inner); cudaFree(s->d_counters);
    for (cudaEvent_t ev : s->counter_event) if (ev) cudaEventDestroy(ev);
    if (s->h_scratch) cudaFreeHost(s->h_scratch);
    if (s->d_scratch) cudaFree(s->d_scratch);
    if (s->scratch_stream) cudaStreamDestroy(s->scratch_stream);
    delete s;
    return LSCPM_OK;
}

int lscpm_scene_n_bins(const LscpmScene* s) { return s ? s->dev.n_bins : fail(LSCPM_EINVAL, "scene is NULL"); }

int lscpm_scene_kernel_name(const LscpmScene* s, char* buf, int buf_len) {
    if (!s) return fail(LSCPM_EINVAL, "scene is NULL");
    char text[160];
    if (s->pool) {
        snprintf(text, sizeof(text), "trace_pool_kernel<ext=%d, flight=%d, warps=%d, records=%d>%s", s->dev.n_ext > 0 ? 1 : 0,
                 s->flight ? 1 : 0, pool_warps(s->pool), pool_slots(s->pool), s->pool_small_plan ? " (records=120 for plans <= 16 KB)" : "");
    } else if (s->sorted) {
        snprintf(text, sizeof(text), "trace_sorted_kernel<ext=%d, flight=%d, block=%d>%s", s->dev.n_ext > 0 ? 1 : 0, s->flight ? 1 : 0, SORT_THREADS,
                 s->sorted_forced ? "" : " (trace_kernel for launches below 4 blocks' worth of photons per resident block)");
    } else {
        snprintf(text, sizeof(text), "trace_kernel<ext=%d, flight=%d, block=%d>", s->dev.n_ext > 0 ? 1 : 0, s->flight ? 1 : 0, s->trace_block);
    }
    if (buf && buf_len > 0) { strncpy(buf, text, (size_t)buf_len - 1); buf[buf_len - 1] = 0; }
    return (int)strlen(text);
}

int lscpm_scene_bin_name(const LscpmScene* s, int bin, char* buf, int buf_len) {
    if (!s) return fail(LSCPM_EINVAL, "scene is NULL");
    if (bin < 0 || bin >= s->dev.n_bins) return fail(LSCPM_EINVAL, "bin index out of range");
    const std::string& name = s->bin_names[bin];
    if (buf && buf_len > 0) { strncpy(buf, name.c_str(), (size_t)buf_len - 1); buf[buf_len - 1] = 0; }
    return (int)name.size();
}

namespace {

inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

int plan_layout(const LscpmBatchDesc* batches, int32_t n_batches, const LscpmSpectraDesc* spectra, PlanLayout& L) {
    if (!batches || n_batches < 1) return fail(LSCPM_EINVAL, "scene/batches missing");
    size_t knots = 0, n_spec = 0;
    if (spectra && spectra->n_spectra > 0) {
        n_spec = (size_t)spectra->n_spectra;
        knots = (size_t)(spectra->begin[n_spec] - spectra->begin[0]);
        if (spectra->begin[0] != 0) return fail(LSCPM_EINVAL, "spectra->begin must start at 0");
    }
    L.n_knots = knots; L.n_guide = n_spec * (SPEC_BUCKETS + 1);
    L.off_batches = 0;
    L.off_x = align16((size_t)n_batches * sizeof(DevBatch));
    L.off_cdf = L.off_x + align16(std::max<size_t>(knots, 1) * sizeof(float));
    L.off_guide = L.off_cdf + align16(std::max<size_t>(knots, 1) * sizeof(float));
    L.bytes = L.off_guide + align16(std::max<size_t>(L.n_guide, 1) * sizeof(unsigned short));
    return LSCPM_OK;
}

// lower batches and spectra into `image` (L.bytes bytes of host memory, pinned or not)
int plan_fill(const LscpmScene* s, const LscpmBatchDesc* batches, int32_t n_batches, const LscpmSpectraDesc* spectra,
              const PlanLayout& L, unsigned char* image) {
    DevBatch* lowered = reinterpret_cast<DevBatch*>(image + L.off_batches);
    float* sx = reinterpret_cast<float*>(image + L.off_x);
    float* sc = reinterpret_cast<float*>(image + L.off_cdf);
    unsigned short* sg = reinterpret_cast<unsigned short*>(image + L.off_guide);
    std::vector<int> sb(1, 0);
    if (spectra && spectra->n_spectra > 0) {
        sb.reserve((size_t)spectra->n_spectra + 1);
        std::vector<double> cdf;
        size_t w = 0, gw = 0;
        for (int i = 0; i < spectra->n_spectra; ++i) {
            const int lo = spectra->begin[i], n = spectra->begin[i + 1] - lo;
            if (n < 2) return fail(LSCPM_EINVAL, "spectrum needs at least two knots");
            if (n > 65000) return fail(LSCPM_EUNSUPPORTED, "spectrum with more than 65000 knots");
            double total = 0;
            for (int k = 0; k < n; ++k) {
                if (!(spectra->y[lo + k] >= 0)) return fail(LSCPM_EINVAL, "spectrum weights must be non-negative");
                if (k && !(spectra->x[lo + k] > spectra->x[lo + k - 1])) return fail(LSCPM_EINVAL, "spectrum wavelengths must increase");
                total += spectra->y[lo + k];
            }
            if (!(total > 0)) return fail(LSCPM_EINVAL, "spectrum is identically zero");
            cumulative_cdf(spectra->x + lo, spectra->y + lo, n, cdf);
            for (int k = 0; k < n; ++k) { sx[w] = (float)spectra->x[lo + k]; sc[w] = (float)cdf[k]; ++w; }
            sb.push_back((int)w);
            int j = 0;
            for (int b = 0; b <= SPEC_BUCKETS; ++b) {
                const double level = (double)b / SPEC_BUCKETS;
                while (j + 2 < n && cdf[j + 1] <= level) ++j;
                sg[gw++] = (unsigned short)j;
            }
        }
    }
    for (int b = 0; b < n_batches; ++b) {
        const int rc = lower_batch(s, batches + b, sb, lowered[b]);
        if (rc) return rc;
    }
    return LSCPM_OK;
}

void plan_point(LscpmPlan* p, unsigned char* d_blob, const PlanLayout& L) {
    p->d_blob = d_blob;
    p->d_batches = reinterpret_cast<DevBatch*>(d_blob + L.off_batches);
    p->d_spec_x = reinterpret_cast<float*>(d_blob + L.off_x);
    p->d_spec_cdf = reinterpret_cast<float*>(d_blob + L.off_cdf);
    p->d_spec_guide = reinterpret_cast<unsigned short*>(d_blob + L.off_guide);
    p->h2d_bytes = L.bytes;
}

// the scene's scratch buffers, grown geometrically; caller holds scratch_mutex
int scratch_reserve(LscpmScene* s, size_t host_bytes, size_t dev_bytes) {
    if (!s->scratch_stream) CUDA_TRY(cudaStreamCreateWithFlags(&s->scratch_stream, cudaStreamNonBlocking));
    if (host_bytes > s->h_scratch_bytes) {
        if (s->h_scratch) cudaFreeHost(s->h_scratch);
        s->h_scratch = nullptr; s->h_scratch_bytes = 0;
        const size_t want = std::max<size_t>(host_bytes + host_bytes / 2, 64 * 1024);
        CUDA_TRY(cudaMallocHost((void**)&s->h_scratch, want));
        s->h_scratch_bytes = want;
    }
    if (dev_bytes > s->d_scratch_bytes) {
        if (s->d_scratch) cudaFree(s->d_scratch);
        s->d_scratch = nullptr; s->d_scratch_bytes = 0;
        const size_t want = std::max<size_t>(dev_bytes + dev_bytes / 2, 64 * 1024);
        CUDA_TRY(cudaMalloc((void**)&s->d_scratch, want));
        s->d_scratch_bytes = want;
    }
    return LSCPM_OK;
}

}  // namespace

// One device allocation and one host->device copy per plan: the lowered descriptors and the spectra are written straight
// into the scene's pinned staging buffer and uploaded with a single cudaMemcpyAsync.
int lscpm_plan_create(LscpmScene* s, const LscpmBatchDesc* batches, int32_t n_batches, const LscpmSpectraDesc* spectra,
                      LscpmPlan** out) {
    if (!out) return fail(LSCPM_EINVAL, "out pointer is NULL");
    *out = nullptr;
    if (!s || !batches || n_batches < 1) return fail(LSCPM_EINVAL, "scene/batches missing");
    CUDA_TRY(cudaSetDevice(s->device));
    PlanLayout L;
    int rc = plan_layout(batches, n_batches, spectra, L);
    if (rc) return rc;
    LscpmPlan* p = new (std::nothrow) LscpmPlan();
    if (!p) return fail(LSCPM_ENOMEM, "out of host memory");
    p->scene = s; p->n_batches = n_batches;
    unsigned char* d_blob = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_blob, L.bytes);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc plan"); }
    plan_point(p, d_blob, L);
    {
        std::lock_guard<std::mutex> lock(s->scratch_mutex);
        if ((rc = scratch_reserve(s, L.bytes, 0)) || (rc = plan_fill(s, batches, n_batches, spectra, L, s->h_scratch))) {
            lscpm_plan_destroy(p);
            return rc;
        }
        if ((e = cudaMemcpyAsync(d_blob, s->h_scratch, L.bytes, cudaMemcpyHostToDevice, s->scratch_stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(s->scratch_stream)) != cudaSuccess) {
            lscpm_plan_destroy(p);
            return cuda_fail(e, "plan upload");
        }
    }
    *out = p;
    return LSCPM_OK;
}

uint64_t lscpm_plan_h2d_bytes(const LscpmPlan* p) { return p ? p->h2d_bytes : 0; }

int lscpm_plan_destroy(LscpmPlan* p) {
    if (!p) return LSCPM_OK;
    if (p->owns_blob) cudaFree(p->d_blob);
    delete p;
    return LSCPM_OK;
}

int lscpm_trace_plan(LscpmPlan* p, uint64_t photon_begin, uint64_t photon_count, uint64_t seed, int64_t* tallies_dev, void* stream_) {
    if (!p || !tallies_dev) return fail(LSCPM_EINVAL, "plan/tallies missing");
    if (photon_count == 0) return LSCPM_OK;
    LscpmScene* s = p->scene;
    cudaStream_t stream = (cudaStream_t)stream_;
    CUDA_TRY(cudaSetDevice(s->device));
    TraceArgs a{};
    a.batches = p->d_batches; a.spec_x = p->d_spec_x; a.spec_cdf = p->d_spec_cdf; a.spec_guide = p->d_spec_guide;
    a.photon_begin = photon_begin; a.photon_count = photon_count;
    a.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    a.rk = philox_round_keys(a.key);
    a.tallies = (unsigned long long*)tallies_dev;
    const unsigned long long total = photon_count * (unsigned long long)p->n_batches;
    // launches too small to fill the sorted kernel's blocks a few times over go to the lane kernel
    const bool sorted = s->sorted && (s->sorted_forced || total >= 4ULL * (unsigned long long)(s->n_sm * s->sort_blocks_per_sm) * SORT_THREADS);
    const bool pool = !sorted && s->pool != 0;
    const int pool_cfg = (pool && s->pool_small_plan && p->h2d_bytes <= POOL_SMALL_PLAN_BYTES) ? s->pool_small_plan : s->pool;
    const int per_sm = sorted ? s->sort_blocks_per_sm : (pool ? 1 : s->blocks_per_sm);
    const int grid = s->n_sm * per_sm;
    const int lane_warps = pool ? pool_warps(pool_cfg) : s->trace_block / 32;
    // work items: claimed by a block (sorted kernel) or by a warp (lane / pool kernel); ~8 items per claimer, within a batch
    const unsigned long long claimers = sorted ? (unsigned long long)grid : (unsigned long long)grid * (unsigned long long)lane_warps;
    const unsigned long long lo = sorted ? (unsigned long long)SORT_THREADS : 32ULL, hi = sorted ? 16384ULL : 4096ULL;
    unsigned long long chunk = total / (claimers * 8ULL);
    chunk = std::min<unsigned long long>(std::max<unsigned long long>(chunk, lo), hi);
    chunk = std::min<unsigned long long>((chunk + 31ULL) / 32ULL * 32ULL, std::max<unsigned long long>(photon_count, 1ULL));
    a.chunk = (unsigned)chunk;
    a.chunks_per_batch = (photon_count + chunk - 1) / chunk;
    a.n_chunks = a.chunks_per_batch * (unsigned long long)p->n_batches;
    unsigned counter_slot = 0;
    int rc = pick_counter(s, stream, &a.work_counter, &counter_slot);
    if (rc) return rc;
    if (sorted) {
        // a block drinks from two work items at a time; tiny launches need fewer blocks than the grid
        const unsigned long long want = (total + SORT_THREADS - 1) / SORT_THREADS;
        const int blocks = (int)std::max<unsigned long long>(1ULL, std::min<unsigned long long>((unsigned long long)grid, want));
        sorted_kernel(s)<<<blocks, SORT_THREADS, s->sort_smem, stream>>>(s->dev, a);
    } else {
        const unsigned long long warps = (unsigned long long)lane_warps;
        const int blocks = (int)std::min<unsigned long long>((unsigned long long)grid, (a.n_chunks + warps - 1) / warps);
        if (pool) pool_kernel(s, pool_cfg)<<<std::max(blocks, 1), lane_warps * 32, pool_smem(s, pool_cfg), stream>>>(s->dev, a);
        else lane_kernel(s)<<<std::max(blocks, 1), s->trace_block, s->trace_smem, stream>>>(s->dev, a);
    }
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    if ((rc = release_counter(s, stream, counter_slot))) return rc;
#ifdef SORT_TIMING
    if (sorted) {
        unsigned long long t[64];
        cudaStreamSynchronize(stream);
        cudaMemcpyFromSymbol(t, g_sort_timing, sizeof(t));
        fprintf(stderr, "[sort timing] warp-iterations %llu: work %.0f wait %.0f sort %.0f cycles/iteration\n", t[3],
                (double)t[0] / t[3], (double)t[1] / t[3], (double)t[2] / t[3]);
        const char* names[] = {"flat", "cyl", "cell", "absorb", "fresh", "out", "idle"};
        for (int c = 0; c < N_CLS; ++c)
            if (t[9 + 2 * c]) fprintf(stderr, "  pure %-6s warps: %5.1f%% of warp-iterations, %.0f cycles\n", names[c], 100.0 * t[9 + 2 * c] / t[3], (double)t[8 + 2 * c] / t[9 + 2 * c]);
        if (t[41]) fprintf(stderr, "  mixed       warps: %5.1f%% of warp-iterations, %.0f cycles\n", 100.0 * t[41] / t[3], (double)t[40] / t[41]);
        memset(t, 0, sizeof(t));
        cudaMemcpyToSymbol(g_sort_timing, t, sizeof(t));
    }
#endif
    return LSCPM_OK;
}

int lscpm_trace(LscpmScene* s, const LscpmBatchDesc* batches, int32_t n_batches, const LscpmSpectraDesc* spectra,
                uint64_t photon_begin, uint64_t photon_count, uint64_t seed, int64_t* tallies_dev, void* stream) {
    LscpmPlan* plan = nullptr;
    int rc = lscpm_plan_create(s, batches, n_batches, spectra, &plan);
    if (rc) return rc;
    rc = lscpm_trace_plan(plan, photon_begin, photon_count, seed, tallies_dev, stream);
    // the launch reads the plan's device buffers: wait for it before releasing them
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    lscpm_plan_destroy(plan);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "trace kernel");
    return LSCPM_OK;
}

// Host buffers in, host tallies out, no allocation per call: descriptors and spectra are lowered into the scene's pinned
// staging buffer, one copy takes them to the scene's device scratch, the tally block behind them is zeroed, traced into
// and copied back - all on the scene's own stream, one synchronisation at the end.
int lscpm_trace_host(LscpmScene* s, const LscpmBatchDesc* batches, int32_t n_batches, const LscpmSpectraDesc* spectra,
                     uint64_t photon_begin, uint64_t photon_count, uint64_t seed, int64_t* tallies_host) {
    if (!s || !tallies_host) return fail(LSCPM_EINVAL, "scene/tallies missing");
    CUDA_TRY(cudaSetDevice(s->device));
    PlanLayout L;
    int rc = plan_layout(batches, n_batches, spectra, L);
    if (rc) return rc;
    const size_t tally_bytes = (size_t)n_batches * (size_t)(s->dev.n_bins + LSCPM_N_COUNTERS) * sizeof(int64_t);
    std::lock_guard<std::mutex> lock(s->scratch_mutex);
    if ((rc = scratch_reserve(s, std::max(L.bytes, tally_bytes), L.bytes + tally_bytes))) return rc;
    if ((rc = plan_fill(s, batches, n_batches, spectra, L, s->h_scratch))) return rc;
    LscpmPlan plan;
    plan.scene = s; plan.n_batches = n_batches; plan.owns_blob = false;
    plan_point(&plan, s->d_scratch, L);
    int64_t* d_tallies = reinterpret_cast<int64_t*>(s->d_scratch + L.bytes);
    CUDA_TRY(cudaMemcpyAsync(s->d_scratch, s->h_scratch, L.bytes, cudaMemcpyHostToDevice, s->scratch_stream));
    CUDA_TRY(cudaMemsetAsync(d_tallies, 0, tally_bytes, s->scratch_stream));
    if ((rc = lscpm_trace_plan(&plan, photon_begin, photon_count, seed, d_tallies, s->scratch_stream))) return rc;
    // the staging buffer is free again once the upload has been consumed; stream order guarantees that here
    CUDA_TRY(cudaMemcpyAsync(s->h_scratch, d_tallies, tally_bytes, cudaMemcpyDeviceToHost, s->scratch_stream));
    CUDA_TRY(cudaStreamSynchronize(s->scratch_stream));
    memcpy(tallies_host, s->h_scratch, tally_bytes);
    return LSCPM_OK;
}

// ---- per-ray entry points ------------------------------------------------------------------------

namespace {
// pvtrace's geometric container test in double: a body contains the ray origin iff the ray has exactly one
// forward intersection with it beyond the zero tolerance.
constexpr double ZERO_TOL = 1e-12;
int forward_hits_cylinder(double ox, double oz, double dx, double dz, double r, double y, double dy, double hy) {
    // side roots clipped to |y| <= hy, plus caps inside the radius
    int n = 0;
    const double a = dx * dx + dz * dz, b = 2 * (ox * dx + oz * dz), c = ox * ox + oz * oz - r * r;
    if (a > 0) {
        const double disc = b * b - 4 * a * c;
        if (disc >= 0) {
            const double sq = std::sqrt(disc), q = -0.5 * (b + std::copysign(sq, b));
            double t1 = q / a, t2 = q != 0 ? c / q : 0.0;
            for (double t : {t1, t2})
                if (t > ZERO_TOL && std::fabs(y + t * dy) <= hy) ++n;
        }
    }
    if (dy != 0)
        for (double yc : {-hy, hy}) {
            const double t = (yc - y) / dy;
            if (t > ZERO_TOL) { const double x = ox + t * dx, z = oz + t * dz; if (x * x + z * z < r * r) ++n; }
        }
    return n;
}
int forward_hits_box(const double* h, const double* o, const double* d) {
    double tmin = -INFINITY, tmax = INFINITY;
    for (int ax = 0; ax < 3; ++ax) {
        if (d[ax] == 0) { if (std::fabs(o[ax]) > h[ax]) return 0; continue; }
        double lo = (-h[ax] - o[ax]) / d[ax], hi = (h[ax] - o[ax]) / d[ax];
        if (lo > hi) std::swap(lo, hi);
        tmin = std::max(tmin, lo); tmax = std::min(tmax, hi);
    }
    if (tmin > tmax) return 0;
    return (tmin > ZERO_TOL) + (tmax > ZERO_TOL);
}

// plate-local double ray -> medium + capillary-relative float coordinates
void classify_ray(const LscpmScene* s, const double* o, const double* d, ProbeIn& q) {
    int k = s->n_ch > 0 ? (int)std::lrint((o[0] - s->cx0) / s->pitch) : 0;
    k = std::min(std::max(k, 0), std::max(s->n_ch - 1, 0));
    const double xr = s->n_ch > 0 ? o[0] - (s->cx0 + k * s->pitch) : o[0], oz = s->n_ch > 0 ? o[2] - s->zc : o[2];
    q.kc = k; q.xr = (float)xr; q.y = (float)o[1]; q.z = (float)o[2];
    q.dx = (float)d[0]; q.dy = (float)d[1]; q.dz = (float)d[2];
    const double h[3] = {s->hx, s->hy, s->hz};
    q.medium = MED_AIR;
    if (forward_hits_box(h, o, d) == 1) {
        q.medium = MED_PLATE;
        if (s->n_ch > 0 && forward_hits_cylinder(xr, oz, d[0], d[2], s->r_o, o[1], d[1], s->hy) == 1) {
            q.medium = MED_OUTER;
            if (s->has_inner && forward_hits_cylinder(xr, oz, d[0], d[2], s->r_i, o[1], d[1], s->hy) == 1) q.medium = MED_INNER;
        }
    }
}
int node_of(const LscpmScene* s, int medium, int ch) {
    if (medium == MED_AIR) return 0;
    if (medium == MED_PLATE) return s->node_plate;
    if (medium == MED_OUTER) return s->node_outer[ch];
    return s->node_inner[ch];
}
}  // namespace

int lscpm_probe(LscpmScene* s, int64_t n, const double* pos, const double* dir, int32_t* hit, int32_t* container,
                int32_t* adjacent, int32_t* face, double* distance, double* point, double* normal, double* reflectivity,
                double* reflected, double* refracted) {
    if (!s || n < 0 || !pos || !dir) return fail(LSCPM_EINVAL, "scene/rays missing");
    if (n == 0) return LSCPM_OK;
    CUDA_TRY(cudaSetDevice(s->device));
    std::vector<ProbeIn> in((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        double o[3], d[3];
        world_point_to_plate(s->plate_pose, pos + 3 * i, o);
        world_vec_to_plate(s->plate_pose, dir + 3 * i, d);
        classify_ray(s, o, d, in[(size_t)i]);
    }
    ProbeIn* d_in = nullptr; ProbeOut* d_out = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_in, sizeof(ProbeIn) * (size_t)n));
    cudaError_t e = cudaMalloc((void**)&d_out, sizeof(ProbeOut) * (size_t)n);
    if (e != cudaSuccess) { cudaFree(d_in); return cuda_fail(e, "cudaMalloc"); }
    std::vector<ProbeOut> out((size_t)n);
    e = cudaMemcpy(d_in, in.data(), sizeof(ProbeIn) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        probe_kernel<<<(unsigned)((n + 127) / 128), 128>>>(s->dev, d_in, d_out, n);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out.data(), d_out, sizeof(ProbeOut) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "probe kernel");
    for (int64_t i = 0; i < n; ++i) {
        const ProbeOut& o = out[(size_t)i];
        const ProbeIn& q = in[(size_t)i];
        hit[i] = container[i] = adjacent[i] = face[i] = -1;
        distance[i] = reflectivity[i] = 0;
        for (int c = 0; c < 3; ++c) point[3 * i + c] = normal[3 * i + c] = reflected[3 * i + c] = refracted[3 * i + c] = 0;
        if (!o.found) {
            // nothing of the plate assembly ahead: pvtrace reports the world sphere (hit node 0, no adjacent)
            hit[i] = 0; container[i] = 0; adjacent[i] = -1; face[i] = 0;
            const double* pw = pos + 3 * i; const double* dw = dir + 3 * i;
            const double oc[3] = {pw[0] - s->world_centre[0], pw[1] - s->world_centre[1], pw[2] - s->world_centre[2]};
            const double b = oc[0] * dw[0] + oc[1] * dw[1] + oc[2] * dw[2];
            const double c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - s->world_radius * s->world_radius;
            const double disc = b * b - c;
            if (disc < 0 || -b + std::sqrt(disc) <= 0) { hit[i] = container[i] = -1; continue; }   // outside the world
            distance[i] = -b + std::sqrt(disc);
            for (int k = 0; k < 3; ++k) point[3 * i + k] = pw[k] + distance[i] * dw[k];
            continue;
        }
        // node bookkeeping of pvtrace's next_hit for this interface
        int h_node, c_node, a_node;
        if (q.medium == MED_AIR) {
            h_node = o.ext > 0 ? s->node_ext[o.ext - 1] : s->node_plate;
            if (o.found == 2) { c_node = h_node; a_node = 0; }      // already inside that box (zero-distance entry dropped)
            else { c_node = 0; a_node = h_node; }
        }
        else if (o.kind == SURF_BOX) { h_node = c_node = s->node_plate; a_node = 0; }   // beyond the plate: the world
        else if (o.kind == SURF_CAP) {
            // the cap coincides with the plate's face; in scene-graph order the plate comes first (hit node and nearest
            // node left once), the tube second
            h_node = c_node = s->node_plate; a_node = s->node_outer[q.kc];
        }
        else if (o.kind == SURF_OUTER) {
            if (q.medium == MED_PLATE) { h_node = a_node = s->node_outer[o.ch]; c_node = s->node_plate; }
            else { h_node = c_node = s->node_outer[q.kc]; a_node = s->node_plate; }
        } else {
            if (q.medium == MED_OUTER) { h_node = a_node = s->node_inner[q.kc]; c_node = o.container == MED_PLATE ? s->node_plate : s->node_outer[q.kc]; }
            else { h_node = c_node = s->node_inner[q.kc]; a_node = o.adj == MED_PLATE ? s->node_plate : s->node_outer[q.kc]; }
        }
        hit[i] = h_node; container[i] = c_node; adjacent[i] = a_node;
        face[i] = (o.kind == SURF_BOX || o.kind == SURF_CAP) ? o.face : 0;
        distance[i] = o.t;
        for (int c = 0; c < 3; ++c) point[3 * i + c] = pos[3 * i + c] + (double)o.t * dir[3 * i + c];
        const double nl[3] = {o.nx, o.ny, o.nz}, rl[3] = {o.rx, o.ry, o.rz}, tl[3] = {o.tx, o.ty, o.tz};
        plate_vec_to_world(s->plate_pose, nl, normal + 3 * i);
        plate_vec_to_world(s->plate_pose, rl, reflected + 3 * i);
        plate_vec_to_world(s->plate_pose, tl, refracted + 3 * i);
        reflectivity[i] = o.R;
    }
    return LSCPM_OK;
}

int lscpm_emit(LscpmScene* s, const LscpmBatchDesc* batch, const LscpmSpectraDesc* spectra, uint64_t photon_begin,
               int64_t n, uint64_t seed, double* pos, double* dir, double* wavelength) {
    if (!s || !batch || n < 0) return fail(LSCPM_EINVAL, "scene/batch missing");
    if (n == 0) return LSCPM_OK;
    LscpmPlan* plan = nullptr;
    int rc = lscpm_plan_create(s, batch, 1, spectra, &plan);
    if (rc) return rc;
    float* d_out = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_out, sizeof(float) * 7 * (size_t)n);
    std::vector<float> out((size_t)n * 7);
    if (e == cudaSuccess) {
        emit_kernel<<<(unsigned)((n + 127) / 128), 128>>>(s->dev, plan->d_batches, plan->d_spec_x, plan->d_spec_cdf, plan->d_spec_guide, photon_begin,
                                                           n, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), d_out);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out.data(), d_out, sizeof(float) * 7 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    lscpm_plan_destroy(plan);
    if (e != cudaSuccess) return cuda_fail(e, "emit kernel");
    for (int64_t i = 0; i < n; ++i) {
        const double o[3] = {out[7 * i], out[7 * i + 1], out[7 * i + 2]}, d[3] = {out[7 * i + 3], out[7 * i + 4], out[7 * i + 5]};
        plate_point_to_world(s->plate_pose, o, pos + 3 * i);
        plate_vec_to_world(s->plate_pose, d, dir + 3 * i);
        wavelength[i] = out[7 * i + 6];
    }
    return LSCPM_OK;
}

int lscpm_follow(LscpmScene* s, int64_t n, const double* pos, const double* dir, const double* wavelength, uint32_t batch_id,
                 uint64_t photon_begin, uint64_t seed, int32_t max_steps, int32_t* n_steps, int32_t* step_event,
                 int32_t* step_node, double* step_pos, double* step_dir, double* step_wavelength, int32_t* final_bin) {
    if (!s || n < 0 || !pos || !dir || !wavelength || !n_steps || !final_bin || max_steps < 0) return fail(LSCPM_EINVAL, "bad follow arguments");
    if (n == 0) return LSCPM_OK;
    CUDA_TRY(cudaSetDevice(s->device));
    std::vector<FollowIn> in((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        double o[3], d[3];
        world_point_to_plate(s->plate_pose, pos + 3 * i, o);
        world_vec_to_plate(s->plate_pose, dir + 3 * i, d);
        in[(size_t)i] = {(float)o[0], (float)o[1], (float)o[2], (float)d[0], (float)d[1], (float)d[2], (float)wavelength[i]};
    }
    const int ms = std::max(max_steps, 1);
    FollowIn* d_in = nullptr; FollowStep* d_steps = nullptr; int *d_n = nullptr, *d_bin = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_in, sizeof(FollowIn) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_steps, sizeof(FollowStep) * (size_t)n * ms);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_n, sizeof(int) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_bin, sizeof(int) * (size_t)n);
    std::vector<FollowStep> steps((size_t)n * ms);
    std::vector<int> ns((size_t)n), bins((size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, in.data(), sizeof(FollowIn) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_steps, 0, sizeof(FollowStep) * (size_t)n * ms);
    if (e == cudaSuccess) {
        follow_kernel<<<(unsigned)((n + 127) / 128), 128>>>(s->dev, d_in, n, batch_id, photon_begin,
                                                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), max_steps, d_n, d_steps, d_bin);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(steps.data(), d_steps, sizeof(FollowStep) * steps.size(), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ns.data(), d_n, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(bins.data(), d_bin, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_steps); cudaFree(d_n); cudaFree(d_bin);
    if (e != cudaSuccess) return cuda_fail(e, "follow kernel");
    for (int64_t i = 0; i < n; ++i) {
        n_steps[i] = ns[(size_t)i];
        final_bin[i] = bins[(size_t)i];
        if (!step_event) continue;
        for (int k = 0; k < std::min(ns[(size_t)i], max_steps); ++k) {
            const FollowStep& st = steps[(size_t)i * ms + k];
            const size_t at = (size_t)i * max_steps + k;
            step_event[at] = st.event;
            if (step_node) {
                // node of the event: the surface's node for surface events, the container for absorption events
                int node = -1;
                if (st.medium == MED_AIR && !s->node_ext.empty() && st.event != EV_EXIT && st.event != EV_KILL) {
                    node = st.ch == 0 ? s->node_plate : s->node_ext[st.ch - 1];
                } else if (st.event == EV_REFLECT || st.event == EV_TRANSMIT) {
                    if (st.kind == SURF_BOX || st.kind == SURF_CAP) node = s->node_plate;
                    else if (st.kind == SURF_OUTER) node = s->node_outer[st.ch];
                    else if (st.kind == SURF_INNER) node = s->node_inner[st.ch];
                } else if (st.event == EV_ABSORB || st.event == EV_REACT || st.event == EV_EMIT) node = node_of(s, st.medium, st.ch);
                step_node[at] = node;
            }
            if (step_pos) { const double l[3] = {st.x, st.y, st.z}; plate_point_to_world(s->plate_pose, l, step_pos + 3 * at); }
            if (step_dir) { const double l[3] = {st.dx, st.dy, st.dz}; plate_vec_to_world(s->plate_pose, l, step_dir + 3 * at); }
            if (step_wavelength) step_wavelength[at] = st.wl;
        }
    }
    return LSCPM_OK;
}

}  // extern "C"
