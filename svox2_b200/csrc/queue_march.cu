// Sample-queue fused march: the training-step renderer for small ray batches.
//
// Why: with 5000 rays per step a warp-per-ray kernel is ONE wave of 5000 warps whose duration is
// the longest ray; as short rays retire the SMs drain (ncu: 22 % achieved occupancy for the
// lattice kernel).  The expensive work -- gathering 8 SH rows per active sample and scattering
// their gradients -- has no sequential dependence once each sample's compositing weight is
// known, so it is taken out of the per-ray warp and run SAMPLE-parallel over a compact queue:
//
//   K1 trace   (warp per ray)    lattice traversal, sigma, log-transmittance chain (bit-exact,
//                                same code path as march.cu); every active sample becomes a
//                                96-byte record {8 links, pos, weight, logT, sigma, ws, ray};
//                                blocks of <=32 records are bump-allocated with one atomic per
//                                batch and chained per ray.
//   K2 shade   (8 lanes/sample)  128-bit gathers of the 8 corner rows, colours, atomicAdd of
//                                weight*max(colour,0) into rgb_out; colour sums kept in the record.
//   K3 seed    (warp per ray)    MSE/beta seed from the ray's final colour, then the reference's
//                                accum recursion over the ray's records (a warp scan per block);
//                                leaves dL/dcolour and dL/dsigma in each record.
//   K4 scatter (8 lanes/run)     each 8-lane group walks a contiguous run of records, merges
//                                consecutive samples that share a cell, then red.global.add.v4.f32
//                                into grad_sh, density atomics, mask bytes.
//
// K2 and K4 are load balanced over the queue (equal record counts per group) and expose one
// independent gather/scatter per 8 lanes.  If the queue overflows its workspace a device flag makes K2-K4
// no-ops and the single-kernel lattice path (march.cu) redoes the batch, so correctness never
// depends on the workspace size.
//
// Parity: K1 is the reference's traversal (render_lerp_kernel_cuvol.cu:56-109) -> identical sample
// set and mask; colours/gradients are reassociated sums (float atomics), within 1e-6 relative.
#include "common.cuh"
#include "march.cuh"
#include "slots.cuh"
#include "traverse.cuh"

namespace svox2b {
namespace {

// Tunables (A/B on B200, profiles/r1_ab_variants.txt)
#ifndef QM_SAMPLE_CTAS
#define QM_SAMPLE_CTAS 16  // persistent CTAs per SM for the sample-parallel kernels
#endif

#ifndef QM_GATHER_AHEAD
#define QM_GATHER_AHEAD 1  // K2 prefetches the rows of the group's NEXT record to L2
#endif
#ifndef QM_TRACE_BLOCKS
#define QM_TRACE_BLOCKS 16
#endif
#ifndef QM_SCATTER_BLOCKS
// min resident CTAs per SM asked of ptxas for the slot-scatter variant of K4 (register cap = 65536 /
// (256 * n)).  A/B on B200 (profiles/r2_ab_variants.txt): 4 CTAs (64 registers, 92 B of spills) makes
// the fused step 0.700 -> 0.686 ms; the dense-gradient variant loses with the same cap (fused call
// 0.376 -> 0.400 ms) and keeps 80 registers.
#define QM_SCATTER_BLOCKS 4
#endif
#ifndef QM_DENSE_SCATTER_BLOCKS
#define QM_DENSE_SCATTER_BLOCKS 3
#endif

constexpr int BLK_TABLE = 64;      // blocks per ray kept in the table (later ones: chain only)
constexpr int TRACE_THREADS = 64;
constexpr int SAMPLE_THREADS = 256;

struct __align__(16) Record {
    int32_t lk[8];       // corner links
    float px, py, pz, w; // fractional position, compositing weight
    float after, sigma, ws;
    int32_t ray;
    float c0, c1, c2, cgs;  // K2: colour sums (+0.5); K3: dL/dcolour (weight applied), dL/dsigma
    int32_t blk_count, blk_next, pad0, pad1;  // valid in the first record of a block
};
static_assert(sizeof(Record) == 96, "record layout");

struct QueueView {
    Record* __restrict__ rec;
    int capacity;
    int* __restrict__ counter;    // records allocated so far
    int* __restrict__ overflow;   // set when the queue ran out of space
    int* __restrict__ sticky;     // batches that overflowed before the current one (slab mode)
    int* __restrict__ ray_first;  // [Q] first block of each ray or -1
    float* __restrict__ ray_logT; // [Q] forward log-transmittance (-1e3 when stopped)
    float4* __restrict__ ray_Y;   // [Q][8] SH basis value of channels 4q..4q+3 (0 for padding)
    float* __restrict__ ray_accum;// [Q] what is left of the reference's accum behind the foreground (MSI)
    // per-ray block table for K3: (first record, count) of the ray's first BLK_TABLE blocks and the
    // number of blocks, so that K3 need not walk the chain header by header (NULL: ship mode)
    int2* __restrict__ blk_table; // [Q][BLK_TABLE]
    int* __restrict__ ray_nblk;   // [Q]
    // ship mode (record shipping between slabs): the record array is n_regions sub-arrays of
    // region_cap records, region k filled by rank k; counts[k] = records in region k this step
    const int* __restrict__ counts;
    int n_regions, region_cap;
};

// flat sample index -> record, across the regions of a ship-mode queue.  The prefix table lives in
// shared memory; a thread's indices only grow, so it keeps a running region cursor.
struct Regions {
    int n, cap, total;
    const int* pre;   // [n + 1] records in the regions before r (shared memory)
};
template <int MODE>
__device__ __forceinline__ void regions_init(const QueueView& qv, Regions& rg, int* s_pre) {
    rg.n = 0;
    rg.cap = 0;
    rg.pre = s_pre;
    if (MODE == 2) {
        rg.n = qv.n_regions;
        rg.cap = qv.region_cap;
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int r = 0; r < rg.n; ++r) {
                s_pre[r] = acc;
                acc += min(__ldcg(qv.counts + r), rg.cap);
            }
            s_pre[rg.n] = acc;
        }
        __syncthreads();
        rg.total = s_pre[rg.n];
    } else {
        rg.total = min(*qv.counter, qv.capacity);
    }
}
template <int MODE>
__device__ __forceinline__ int region_index(const Regions& rg, int s, int& cursor) {
    if (MODE != 2) return s;
    while (cursor + 1 < rg.n && s >= rg.pre[cursor + 1]) ++cursor;
    return cursor * rg.cap + (s - rg.pre[cursor]);
}

// Ship mode (slab.py): rays are sharded over the ranks for the TRAVERSAL, the grid for everything
// else.  The rank that traces a ray composites it exactly as the single-GPU kernel does and ships
// every active sample's record to the rank that owns the sample's cell, straight into region
// `rank` of that rank's queue over NVLink (peer-mapped symmetric memory).
struct ShipView {
    Record* rec[SVOX2_MAX_PEERS];       // peer d's record array, already offset to MY region
    int* ray_first[SVOX2_MAX_PEERS];    // peer d's [Qg] chain heads
    float* ray_logT[SVOX2_MAX_PEERS];   // peer d's [Qg] forward log-transmittance
    int* overflow[SVOX2_MAX_PEERS];     // peer d's overflow flag
    int bounds[SVOX2_MAX_PEERS + 1];    // slab x boundaries
    int* cnt;                           // local [n]: records shipped to each peer so far
    int n, region_base, region_cap;
    int64_t ray_offset;                 // global id of this rank's first ray
};

// Slab mode (grid sharded over GPUs by x-range of cells, slab.py): the ray is traversed in full so
// that the sample lattice is the reference's, but only samples whose cell lies in [x_lo, x_hi)
// are gathered / recorded; transmittance is accumulated from 0 and shifted by the other slabs'
// contribution (lt_in) once that is known.
struct SlabView {
    int x_lo, x_hi;
    float* __restrict__ logT_local;        // K1 out: [Q] sum of -sigma*delta over this slab's samples
    const float* __restrict__ lt_in;       // K2-K4 in: [Q] log-transmittance at slab entry
    const float* __restrict__ gout;        // K3 in: [Q,3] dL/d(rgb) of the whole ray
    const float* __restrict__ accum_start; // K3 in: [Q] the reference's accum at slab entry
    float* __restrict__ stopped;           // K2 out: stride 4 (the .w of the [Q,4] partials): 1 where the stop fired
    float stop_thresh;
};


// ---------------------------------------------------------------------------------------
// K1: traversal -> records
// ---------------------------------------------------------------------------------------
template <bool COUNT, int MODE>
__global__ void __launch_bounds__(TRACE_THREADS, QM_TRACE_BLOCKS)
trace_kernel(GridView g, Opts opt, MarchArgs a, QueueView qv, SlabView sl, ShipView sv) {
    constexpr bool SLAB = MODE == 1, SHIP = MODE == 2;
    __shared__ int4 ship_stage[SHIP ? TRACE_THREADS / WARP : 1][SHIP ? WARP * 6 : 1];
    int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray_id >= a.n_rays) return;
    if (SHIP) ray_id += sv.ray_offset;   // a.n_rays = this rank's share; arrays hold the global batch
    const int lane = threadIdx.x & 31;
    float ow[3], dw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ow[i] = __ldg(a.origins + ray_id * 3 + i);
        dw[i] = __ldg(a.dirs + ray_id * 3 + i);
    }
    if (!SHIP && lane < 8) {  // per-ray basis table for the sample-parallel kernels
        float y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = lane * 4 + e;
            y[e] = j < g.sh_dim ? sh_basis_k(g.basis_dim, j % g.basis_dim, dw) : 0.f;
        }
        qv.ray_Y[ray_id * 8 + lane] = make_float4(y[0], y[1], y[2], y[3]);
    }
    Ray ray;
    ray_setup(ray, ow, dw, g, opt);
    const float step = opt.step_size;
    unsigned cnt[3] = {0u, 0u, 0u};
    float lt = 0.f;
    bool stopped = false;
    int first_blk = -1, prev_blk = -1;
    // ship mode: destination of the block whose header is still pending, that block, its size
    int cur_dest = -1, pend_blk = -1, pend_n = 0;
    unsigned visited = 0u;
    int n_blk = 0;            // blocks of this ray so far (non-ship modes)

    // slab mode: t-interval in which the x coordinate can fall into the slab (one cell of margin;
    // the outermost slabs are open-ended because locate() clamps)
    bool takes_part = ray.tmin <= ray.tmax;
    float t_exit = 3.0e38f;
    if (SLAB && takes_part) {
        const float lo = sl.x_lo <= 0 ? -3.0e38f : (float)(sl.x_lo - 1);
        const float hi = sl.x_hi >= g.size[0] - 1 ? 3.0e38f : (float)(sl.x_hi + 1);
        if (ray.d[0] == 0.f) {
            takes_part = ray.o[0] >= lo && ray.o[0] <= hi;
        } else {
            const float ta = (lo - ray.o[0]) * ray.invd[0], tb = (hi - ray.o[0]) * ray.invd[0];
            t_exit = fmaxf(ta, tb);
            takes_part = t_exit >= ray.tmin && fminf(ta, tb) <= ray.tmax;
        }
    }

    if (takes_part) {
        float t0 = ray.tmin;
        bool done = false;
        while (!done) {
            if (SLAB && t0 > t_exit) break;  // past the slab: nothing left for this rank
            // one round = 32 lattice candidates t0 + lane*step (traverse.cuh)
            Candidates c;
            if (SLAB) load_candidates<true>(ray, g, t0, step, lane, c, sl.x_lo, sl.x_hi);
            else load_candidates<false>(ray, g, t0, step, lane, c);
            float t_next = 0.f;
            unsigned n_skips = 0;
            unsigned members = resolve_lattice(ray, c, step, t_next, done, n_skips);
            if (COUNT) cnt[0] += n_skips;
            if (!done) {
                if (SLAB) prefetch_next_links(ray, g, t_next, step, lane, sl.x_lo, sl.x_hi);
                else prefetch_next_links(ray, g, t_next, step, lane);
            }
            if (SLAB) members &= __ballot_sync(FULL, c.l[0] >= sl.x_lo && c.l[0] < sl.x_hi);
            const bool member = (members >> lane) & 1u;
            const float sigma = member_sigma(g, c, member);
            const float tk = c.tk;
            const int32_t (&lk)[8] = c.lk;
            const float (&pos)[3] = c.pos;
            const float ws = (opt.last_sample_opaque && __fadd_rn(tk, step) > ray.tmax) ? 1e9f
                                                                                       : ray.world_step;
            const bool active = member && (sigma > opt.sigma_thresh);
            const float pcnt = __fmul_rn(ws, sigma);
            float my_before;
            unsigned act_m = transmittance_chain(active, pcnt, opt.stop_thresh, lane, lt, members, done,
                                                 stopped, my_before);
            if (COUNT) {
                cnt[1] += __popc(members);
                cnt[2] += __popc(act_m);
            }
            const int n = __popc(act_m);
            if (SHIP) {
                // owner of this lane's cell = slab whose x-range holds it
                int my_dest = 0;
#pragma unroll
                for (int k = 1; k < SVOX2_MAX_PEERS; ++k) my_dest = (k < sv.n && c.l[0] >= sv.bounds[k]) ? k : my_dest;
                unsigned rem = act_m;
                while (rem) {  // samples of a round are in t-order, owners monotone: 1-2 iterations
                    const int lead = __ffs(rem) - 1;
                    const int d = __shfl_sync(FULL, my_dest, lead);
                    const unsigned dm = __ballot_sync(FULL, my_dest == d) & rem;
                    rem &= ~dm;
                    const int nd = __popc(dm);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(sv.cnt + d, nd);
                    base = __shfl_sync(FULL, base, 0);
                    if (base + nd > sv.region_cap) {
                        if (lane == 0) {
                            atomicExch(qv.overflow, 1);
                            *reinterpret_cast<volatile int*>(sv.overflow[d]) = 1;
                        }
                        done = true;
                        break;
                    }
                    // The nd records of this block are contiguous in rank d's queue: stage them in
                    // shared memory and copy them out with consecutive lanes on consecutive 16-byte
                    // chunks, so that they cross NVLink as full 128-byte lines (isolated 16-byte
                    // stores at a 96-byte stride made the kernel twice as slow, r2_ab_variants.txt).
                    // Chunk 5 of record 0 is the block header: written by lane 0 once its successor
                    // is known.
                    int4* const stage = ship_stage[threadIdx.x >> 5];
                    if ((dm >> lane) & 1u) {
                        const int rank = __popc(dm & ((1u << lane) - 1u));
                        int4* r4 = stage + rank * 6;
                        r4[0] = make_int4(lk[0], lk[1], lk[2], lk[3]);
                        r4[1] = make_int4(lk[4], lk[5], lk[6], lk[7]);
                        const float weight = __fmul_rn(fexp(my_before), __fsub_rn(1.f, fexp(-pcnt)));
                        r4[2] = make_int4(__float_as_int(pos[0]), __float_as_int(pos[1]), __float_as_int(pos[2]),
                                          __float_as_int(weight));
                        r4[3] = make_int4(__float_as_int(__fsub_rn(my_before, pcnt)), __float_as_int(sigma),
                                          __float_as_int(ws), (int)ray_id);
                        r4[4] = make_int4(0, 0, 0, 0);
                        r4[5] = make_int4(0, 0, 0, 0);
                    }
                    __syncwarp();
                    {
                        int4* const dst = reinterpret_cast<int4*>(sv.rec[d] + base);
                        for (int ch = lane; ch < nd * 6; ch += WARP)
                            if (ch != 5) dst[ch] = stage[ch];
                    }
                    __syncwarp();
                    if (lane == 0) {
                        // every header / chain head is written exactly once: the header of a block
                        // when its successor (or the end of the ray's stay in that slab) is known
                        const int abs_blk = sv.region_base + base;
                        if (d == cur_dest) {
                            reinterpret_cast<int4*>(sv.rec[d] + pend_blk)[5] = make_int4(pend_n, abs_blk, 0, 0);
                        } else {
                            if (cur_dest >= 0)
                                reinterpret_cast<int4*>(sv.rec[cur_dest] + pend_blk)[5] = make_int4(pend_n, -1, 0, 0);
                            sv.ray_first[d][ray_id] = abs_blk;
                        }
                    }
                    visited |= 1u << d;
                    cur_dest = d;
                    pend_blk = base;
                    pend_n = nd;
                }
            } else if (n > 0) {
                int base = 0;
                if (lane == 0) base = atomicAdd(qv.counter, n);
                base = __shfl_sync(FULL, base, 0);
                if (base + n > qv.capacity) {
                    if (lane == 0) atomicExch(qv.overflow, 1);
                    done = true;  // the fallback path redoes the whole batch of rays
                } else {
                    if ((act_m >> lane) & 1u) {
                        const int rank = __popc(act_m & ((1u << lane) - 1u));
                        Record* r = qv.rec + base + rank;
                        int4* r4 = reinterpret_cast<int4*>(r);
                        r4[0] = make_int4(lk[0], lk[1], lk[2], lk[3]);
                        r4[1] = make_int4(lk[4], lk[5], lk[6], lk[7]);
                        // slab mode keeps the local log-transmittance; K2 makes the weight
                        const float weight = SLAB ? my_before
                                                  : __fmul_rn(fexp(my_before), __fsub_rn(1.f, fexp(-pcnt)));
                        reinterpret_cast<float4*>(r)[2] = make_float4(pos[0], pos[1], pos[2], weight);
                        r4[3] = make_int4(__float_as_int(__fsub_rn(my_before, pcnt)), __float_as_int(sigma),
                                          __float_as_int(ws), (int)ray_id);
                    }
                    if (lane == 0) {  // block header, chain and table: always written by lane 0
                        reinterpret_cast<int4*>(qv.rec + base)[5] = make_int4(n, -1, 0, 0);
                        if (prev_blk >= 0) qv.rec[prev_blk].blk_next = base;
                        else first_blk = base;
                        if (n_blk < BLK_TABLE) qv.blk_table[ray_id * BLK_TABLE + n_blk] = make_int2(base, n);
                    }
                    ++n_blk;
                    prev_blk = base;
                }
            }
            t0 = t_next;
        }
    }
    // forward colour starts from the background term; K2 adds the samples atomically
    const bool hit = ray.tmin <= ray.tmax;
    const float logT_fwd = hit ? (stopped ? -1e3f : lt) : 0.f;
    if (SHIP) {
        if (lane == 0 && cur_dest >= 0)
            reinterpret_cast<int4*>(sv.rec[cur_dest] + pend_blk)[5] = make_int4(pend_n, -1, 0, 0);
        if (lane < sv.n) {  // lane d tells rank d: the ray's transmittance, and "nothing for you"
            sv.ray_logT[lane][ray_id] = logT_fwd;
            if (!((visited >> lane) & 1u)) sv.ray_first[lane][ray_id] = -1;
        }
        if (COUNT && lane == 0) {
            atomicAdd(a.counters + 0, (unsigned long long)cnt[0]);
            atomicAdd(a.counters + 1, (unsigned long long)cnt[1]);
            atomicAdd(a.counters + 2, (unsigned long long)cnt[2]);
        }
        return;
    }
    if (SLAB) {
        if (lane == 0) sl.logT_local[ray_id] = lt;
    } else {
        // with an MSI background the constant-colour term is added by the background kernel from
        // the transmittance left behind the shells (cuvol.cu:1099-1112)
        float bgc = hit ? fexp(logT_fwd) * opt.background_brightness : opt.background_brightness;
        if (g.bg_nlayers > 0) bgc = 0.f;
        if (lane < 3) a.rgb_out[ray_id * 3 + lane] = bgc;
    }
    if (lane == 0) {
        if (!SLAB && !hit) first_blk = -2;  // the seed kernel tells "missed the grid" from "no active sample"
        qv.ray_first[ray_id] = first_blk;
        qv.ray_nblk[ray_id] = n_blk;
        qv.ray_logT[ray_id] = logT_fwd;
        if (COUNT) {
            atomicAdd(a.counters + 0, (unsigned long long)cnt[0]);
            atomicAdd(a.counters + 1, (unsigned long long)cnt[1]);
            atomicAdd(a.counters + 2, (unsigned long long)cnt[2]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// helpers of the sample-parallel kernels
// ---------------------------------------------------------------------------------------
template <bool VEC>
__device__ __forceinline__ float4 load_quad(const float* __restrict__ p, int qi, int D) {
    if (VEC) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v;
    const int j = qi * 4;
    v.x = (j + 0 < D) ? __ldg(p + 0) : 0.f;
    v.y = (j + 1 < D) ? __ldg(p + 1) : 0.f;
    v.z = (j + 2 < D) ? __ldg(p + 2) : 0.f;
    v.w = (j + 3 < D) ? __ldg(p + 3) : 0.f;
    return v;
}
template <bool VEC>
__device__ __forceinline__ void red_quad(float* __restrict__ p, int qi, int D, float a, float b,
                                         float c, float d) {
    if (VEC) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b),
                     "f"(c), "f"(d)
                     : "memory");
    } else {
        const int j = qi * 4;
        if (j + 0 < D) atomicAdd(p + 0, a);
        if (j + 1 < D) atomicAdd(p + 1, b);
        if (j + 2 < D) atomicAdd(p + 2, c);
        if (j + 3 < D) atomicAdd(p + 3, d);
    }
}

// cp.async (LDGSTS) staging of records: the sample-parallel kernels walk records whose addresses are
// known iterations ahead, so the next records travel global -> shared memory while the current one is
// processed, at no register cost (ncu, r2: half of K4's stall samples sat on the first use of the
// record it had just asked for).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#ifndef QM_RING
#define QM_RING 4   // records in flight per 8-lane group (96 bytes of shared memory each)
#endif

__device__ __forceinline__ void corner_weights(float px, float py, float pz, float* wq) {
    const float ax = 1.f - px, ay = 1.f - py, az = 1.f - pz;
    const float yz00 = ay * az, yz01 = ay * pz, yz10 = py * az, yz11 = py * pz;
    wq[0] = yz00 * ax; wq[1] = yz01 * ax; wq[2] = yz10 * ax; wq[3] = yz11 * ax;
    wq[4] = yz00 * px; wq[5] = yz01 * px; wq[6] = yz10 * px; wq[7] = yz11 * px;
}

// ---------------------------------------------------------------------------------------
// K2: gather + colours, 8 lanes per sample
// ---------------------------------------------------------------------------------------
#ifndef QM_SHADE_BLOCKS
// min resident CTAs per SM asked of ptxas: 4 = a 64-register cap.  The variant that also marks the
// touched rows needs 79 registers on its own (3 CTAs per SM); capped it spills 36 bytes and the fused
// step gains 9 us (profiles/r2_ab_variants.txt, sh4).  The plain variant needs 62 either way.
#define QM_SHADE_BLOCKS 4
#endif
template <bool VEC, int MODE, bool MARK>
__global__ void __launch_bounds__(SAMPLE_THREADS, QM_SHADE_BLOCKS)
shade_kernel(GridView g, QueueView qv, float* __restrict__ rgb_out, SlabView sl, SlotMap sm) {
    constexpr bool SLAB = MODE == 1, SHIP = MODE == 2;
    if (*qv.overflow) return;
    __shared__ int s_pre[SVOX2_MAX_PEERS + 1];
    Regions rg;
    regions_init<MODE>(qv, rg, s_pre);
    int cur0 = 0, cur1 = 0;
    const int n = rg.total;
    const int qi = threadIdx.x & 7;
    const int D = g.sh_dim, B = g.basis_dim;
    int ce[4];  // colour of each of this lane's channels (lane constant)
#pragma unroll
    for (int e = 0; e < 4; ++e) {   // channel j belongs to colour j / B (j < 3B: two compares, no division)
        const int j = qi * 4 + e;
        ce[e] = (j < D) ? (j >= B) + (j >= 2 * B) : -1;
    }
    const float* shq = g.sh + (qi * 4 < D ? qi * 4 : 0);
    const unsigned gmask = 0xffu << (threadIdx.x & 24);  // the 8 lanes that share a sample
    const int groups = (gridDim.x * blockDim.x) >> 3;
    // the group's next QM_RING records (s, s + groups, ...) are on their way into shared memory:
    // lane q < 4 copies 16-byte chunk q (links, links, pos + weight, transmittance / sigma / ray)
    __shared__ int4 ring[SAMPLE_THREADS / 8][QM_RING][4];
    int4 (*const my_ring)[4] = ring[threadIdx.x >> 3];
    auto stage = [&](int s, int k) {
        if (s < n && qi < 4)
            cp_async16(&my_ring[k % QM_RING][qi],
                       reinterpret_cast<const int4*>(qv.rec + region_index<MODE>(rg, s, cur1)) + qi);
        cp_async_commit();   // (an empty group keeps the count per iteration uniform)
    };
    const int s_first = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
#pragma unroll
    for (int k = 0; k < QM_RING; ++k) stage(s_first + k * groups, k);
    int k = 0;
    for (int s = s_first; s < n; s += groups, ++k) {
        const int si = region_index<MODE>(rg, s, cur0);
        cp_async_wait<QM_RING - 1>();   // this lane's chunk of record s has landed ...
        __syncwarp(gmask);              // ... and so have the other lanes'
        const int4(&slot)[4] = my_ring[k % QM_RING];
        const int4 la = slot[0];
        const int4 lb = slot[1];
        float4 pw = *reinterpret_cast<const float4*>(&slot[2]);
        const float4 st = *reinterpret_cast<const float4*>(&slot[3]);
        const int ray = __float_as_int(st.w);
#if QM_GATHER_AHEAD & 1
        if (s + groups < n) {  // lane q of the group pulls corner q of the next record towards L2
            cp_async_wait<QM_RING - 2>();
            __syncwarp(gmask);
            const int32_t ln = reinterpret_cast<const int32_t*>(&my_ring[(k + 1) % QM_RING][0])[qi];
            if (ln >= 0) {
                const float* row = g.sh + (size_t)ln * (unsigned)g.sh_stride;
                prefetch_l2(row);
                prefetch_l2(row + (D - 1));
            }
        }
#endif
        __syncwarp(gmask);              // everybody has read slot k: refill it
        stage(s + QM_RING * groups, k + QM_RING);
        if (SLAB) {
            // pw.w is the slab-local log-transmittance before the sample (cuvol.cu:170-178 with the
            // entry value added); a sample behind the stop threshold is dead for K3 / K4 too
            const float lt_in = __ldg(sl.lt_in + ray);
            const float before = __fadd_rn(lt_in, pw.w);
            if (fexp(before) < sl.stop_thresh) {
                if (qi == 0) qv.rec[si].w = -1.f;
                continue;
            }
            const float pcnt = __fmul_rn(st.z, st.y);
            pw.w = __fmul_rn(fexp(before), __fsub_rn(1.f, fexp(-pcnt)));
            if (qi == 0) {
                qv.rec[si].w = pw.w;
                if (fexp(__fadd_rn(lt_in, st.x)) < sl.stop_thresh) sl.stopped[(int64_t)ray * 4] = 1.f;
            }
        }
        const int32_t L[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const unsigned row = L[q] >= 0 ? (unsigned)L[q] : 0u;
            v[q] = load_quad<VEC>(shq + (size_t)row * (unsigned)g.sh_stride, qi, D);
        }
        if (MARK) {  // fused train step: lane q enters corner q into the step's touched-row set
            int32_t Lown = L[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) Lown = (qi == q) ? L[q] : Lown;
            if (Lown >= 0) slot_mark(sm, Lown);
        }
        const float4 Y = __ldcg(qv.ray_Y + (int64_t)ray * 8 + qi);
        float wq[8];
        corner_weights(pw.x, pw.y, pw.z, wq);
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (L[q] >= 0) {
                u.x = fmaf(wq[q], v[q].x, u.x);
                u.y = fmaf(wq[q], v[q].y, u.y);
                u.z = fmaf(wq[q], v[q].z, u.z);
                u.w = fmaf(wq[q], v[q].w, u.w);
            }
        const float pr[4] = {Y.x * u.x, Y.y * u.y, Y.z * u.z, Y.w * u.w};
        float col[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float x = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) x += (ce[e] == c) ? pr[e] : 0.f;
            x += __shfl_xor_sync(gmask, x, 1);
            x += __shfl_xor_sync(gmask, x, 2);
            x += __shfl_xor_sync(gmask, x, 4);
            col[c] = x + 0.5f;
        }
        if (qi < 3) {
            const float cv = qi == 0 ? col[0] : (qi == 1 ? col[1] : col[2]);
            atomicAdd(rgb_out + (int64_t)ray * ((SLAB || SHIP) ? 4 : 3) + qi, pw.w * fmaxf(cv, 0.f));
        }
        if (qi == 0) reinterpret_cast<float4*>(qv.rec + si)[4] = make_float4(col[0], col[1], col[2], 0.f);
    }
}

// ---------------------------------------------------------------------------------------
// K3: gradient seed + accum recursion per ray (cuvol.cu:248-264, 342-351, 735-747)
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TRACE_THREADS)
seed_kernel(Opts opt, MarchArgs a, QueueView qv, SlabView sl) {
    // SLAB: per-ray inputs (dL/drgb, accum at slab entry) come from the exchange, weights and
    // transmittance are relative to the slab entry.  SHIP: the same inputs, but the records carry
    // the tracer's exact weights / transmittance.
    constexpr bool SLAB = MODE != 0, REL = MODE == 1;
    if (*qv.overflow) return;
    const int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray_id >= a.n_rays) return;
    const int lane = threadIdx.x & 31;
    float gout[3], cc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (SLAB) {
            gout[i] = __ldg(sl.gout + ray_id * 3 + i);
            cc[i] = 0.f;
        } else {
            cc[i] = __ldcg(a.rgb_out + ray_id * 3 + i);
            gout[i] = (cc[i] - __ldg(a.grad_in + ray_id * 3 + i)) * a.norm_factor;
        }
    }
    if (!SLAB && a.mse_out != nullptr && lane == 0) {  // F.mse_loss(rgb, rgb_gt) of opt.py:505
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float dlt = cc[i] - __ldg(a.grad_in + ray_id * 3 + i);
            se = fmaf(dlt, dlt, se);
        }
        atomicAdd(a.mse_out, se / (3.f * (float)a.n_rays));
    }
    int blk = qv.ray_first[ray_id];
    float accum = fmaf(cc[0], gout[0], fmaf(cc[1], gout[1], cc[2] * gout[2]));
    float beta_term = 0.f;
    if (!SLAB && a.beta_loss > 0.f && blk != -2) {  // rays that hit the grid (cuvol.cu:259-264)
        const float transmit_in = fexp(qv.ray_logT[ray_id]);
        beta_term = (float)((double)a.beta_loss *
                            (1 - (double)transmit_in / (1 - (double)transmit_in + 1e-3)));
        accum += beta_term;
    }
    if (blk < 0) {  // no active sample: nothing to scatter; MSI hand-over (cuvol.cu:375-377, 1047-1048)
        if (!SLAB && a.accum_out != nullptr && lane == 0) a.accum_out[ray_id] = accum - beta_term;
        return;
    }
    const float lt_in = REL ? __ldg(sl.lt_in + ray_id) : 0.f;
    if (SLAB) accum = __ldg(sl.accum_start + ray_id);  // seed and beta term of the whole ray
    // The ray's blocks in sample order.  With the block table (written by K1) the addresses of all
    // blocks are known up front: block b+1's records are loaded while block b is scanned, instead of
    // one dependent header -> records round trip per block (the kernel was pure L2 latency: 31 us at
    // 8 % issue utilisation).  Blocks behind the table, and ship mode, walk the chain.
    const bool use_table = MODE != 2 && qv.blk_table != nullptr;
    const int n_tab = use_table ? min(__ldcg(qv.ray_nblk + ray_id), BLK_TABLE) : 0;
    int2 ent = make_int2(-1, 0);     // lane j: table entry j of the current window of 32 blocks
    if (n_tab > 0 && lane < n_tab) ent = __ldcg(qv.blk_table + ray_id * BLK_TABLE + lane);
    int tb = 0;                      // next table entry
    // current block: its records in registers
    int cur = blk, cur_count = 0, cur_next = -1;
    float4 pw, st, co;
    auto load_block = [&](int b, int count_hint, int& count, int& next, float4& q_pw, float4& q_st, float4& q_co) {
        // count_hint >= 0: from the table (no header read needed for the addresses)
        count = count_hint >= 0 ? count_hint : __ldcg(&qv.rec[b].blk_count);
        const Record* r = qv.rec + b + (lane < count ? lane : 0);
        q_pw = __ldcg(reinterpret_cast<const float4*>(r) + 2);
        q_st = __ldcg(reinterpret_cast<const float4*>(r) + 3);
        q_co = __ldcg(reinterpret_cast<const float4*>(r) + 4);
        next = -1;
    };
    if (n_tab > 0) {
        cur = __shfl_sync(FULL, ent.x, 0);
        load_block(cur, __shfl_sync(FULL, ent.y, 0), cur_count, cur_next, pw, st, co);
        tb = 1;
    } else {
        load_block(cur, -1, cur_count, cur_next, pw, st, co);
        cur_next = __ldcg(&qv.rec[cur].blk_next);
    }
    while (cur >= 0) {
        // issue the next block's loads before touching this block's data
        int nxt = -1, nxt_count = 0, nxt_next = -1;
        float4 npw = make_float4(0.f, 0.f, 0.f, 0.f), nst = npw, nco = npw;
        if (tb > 0) {   // table mode
            if (tb < n_tab) {
                if ((tb & 31) == 0) ent = (tb + lane < n_tab) ? __ldcg(qv.blk_table + ray_id * BLK_TABLE + tb + lane)
                                                              : make_int2(-1, 0);
                nxt = __shfl_sync(FULL, ent.x, tb & 31);
                load_block(nxt, __shfl_sync(FULL, ent.y, tb & 31), nxt_count, nxt_next, npw, nst, nco);
                ++tb;
            } else if (n_tab == BLK_TABLE) {    // the table is full: the rest of the ray hangs off the chain
                nxt = __ldcg(&qv.rec[cur].blk_next);
                tb = 0;
                if (nxt >= 0) {
                    load_block(nxt, -1, nxt_count, nxt_next, npw, nst, nco);
                    nxt_next = __ldcg(&qv.rec[nxt].blk_next);
                }
            }
        } else {
            nxt = cur_next;
            if (nxt >= 0) {
                load_block(nxt, -1, nxt_count, nxt_next, npw, nst, nco);
                nxt_next = __ldcg(&qv.rec[nxt].blk_next);
            }
        }
        bool on = lane < cur_count;
        Record* r = qv.rec + cur + (on ? lane : 0);
        if (REL) on = on && pw.w >= 0.f;  // dead: behind the stop threshold
        const float w = pw.w, after = REL ? __fadd_rn(lt_in, st.x) : st.x, sigma = st.y, ws = st.z;
        const float col[3] = {co.x, co.y, co.z};
        float gc[3], total = 0.f;
        {
            float tc[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float cl = fmaxf(col[c], 0.f);
                tc[c] = cl * gout[c];
                gc[c] = (cl == col[c]) ? w * gout[c] : 0.f;
            }
            total = tc[0];
            total += tc[2];
            total += tc[1];
        }
        const float wt = on ? w * total : 0.f;
        float pre = wt;  // inclusive scan of weight*total in sample order
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float o = __shfl_up_sync(FULL, pre, off);
            if (lane >= off) pre += o;
        }
        const float my_accum = accum - pre;
        accum -= __shfl_sync(FULL, pre, 31);
        if (on) {
            float cgs = ws * (total * fexp(after) - my_accum);
            if (a.sparsity_loss > 0.f) cgs += a.sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
            reinterpret_cast<float4*>(r)[4] = make_float4(gc[0], gc[1], gc[2], cgs);
        }
        cur = nxt;
        cur_count = nxt_count;
        cur_next = nxt_next;
        pw = npw;
        st = nst;
        co = nco;
    }
    if (!SLAB && a.accum_out != nullptr && lane == 0) a.accum_out[ray_id] = accum - beta_term;
}

// ---------------------------------------------------------------------------------------
// K4: scatter, 8 lanes per contiguous run of records
// ---------------------------------------------------------------------------------------
// Every 8-lane group walks a CONTIGUOUS run of records (consecutive samples of a
// ray lie next to each other in the queue) and keeps the gradient of the current cell in
// registers; consecutive samples that fall into the same cell (~35 % at step 0.5) are merged
// before the REDs leave the SM.  The basis table of the ray is re-read only when the ray changes.
template <bool VEC, int MODE, bool SLOTS>
__global__ void __launch_bounds__(SAMPLE_THREADS, SLOTS ? QM_SCATTER_BLOCKS : QM_DENSE_SCATTER_BLOCKS)
scatter_kernel(GridView g, QueueView qv, GradView gr, int64_t gsh_stride, int64_t gd_stride, SlotMap sm) {
    constexpr bool SLAB = MODE == 1;
    if (*qv.overflow) return;
    __shared__ int s_pre[SVOX2_MAX_PEERS + 1];
    Regions rg;
    regions_init<MODE>(qv, rg, s_pre);
    int cur0 = 0;
    const int n = rg.total;
    const int qi = threadIdx.x & 7;
    const int D = g.sh_dim, B = g.basis_dim;
    int ce[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {   // channel j belongs to colour j / B (j < 3B: two compares, no division)
        const int j = qi * 4 + e;
        ce[e] = (j < D) ? (j >= B) + (j >= 2 * B) : -1;
    }
    const bool quad_on = qi * 4 < D;
    const int groups = (gridDim.x * blockDim.x) >> 3;
    const int per = (n + groups - 1) / groups;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int s0 = min(n, gid * per), s1 = min(n, s0 + per);

    int32_t lown = -1;  // this lane's corner link (corner qi) of the pending cell
    float acc[8][4];
    float dacc = 0.f;
    bool have = false;
    int cur_ray = -1;
    float4 Y = make_float4(0.f, 0.f, 0.f, 0.f);

    int sown = -1;  // SLOTS: slot of this lane's corner of the pending cell
    auto flush = [&]() {
        const unsigned gm = 0xffu << (threadIdx.x & 24);
        const int lane0 = threadIdx.x & 24;
        if (SLOTS) {
            // compact gradient rows: SH and density gradients of a row land in ONE slot row
            float* const base = sm.grad + 4 * qi;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int sq = __shfl_sync(gm, sown, lane0 + q);
                if (quad_on && sq >= 0)
                    red_quad<true>(base + (size_t)(unsigned)sq * (unsigned)sm.width, qi, D, acc[q][0],
                                   acc[q][1], acc[q][2], acc[q][3]);
            }
            if (sown >= 0) {
                atomicAdd(sm.grad + (size_t)(unsigned)sown * (unsigned)sm.width + sm.sigma_at, dacc);
                *reinterpret_cast<uint16_t*>(sm.flags + 2 * (size_t)(unsigned)sown) = 0x0101;  // both masks
            }
        } else {
            // (the eight links of the cell live one per lane: a shuffle per corner instead of eight
            // registers per lane)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int32_t Lq = __shfl_sync(gm, lown, lane0 + q);
                if (quad_on && Lq >= 0)
                    red_quad<VEC>(gr.grad_sh + (int64_t)Lq * gsh_stride + 4 * qi, qi, D, acc[q][0],
                                  acc[q][1], acc[q][2], acc[q][3]);
            }
            if (lown >= 0) {
                atomicAdd(gr.grad_density + (int64_t)lown * gd_stride, dacc);
                if (gr.mask) gr.mask[lown] = 1;
            }
        }
    };

    // records s .. s+QM_RING-1 of this group's run are on their way into shared memory: lane q < 5
    // copies 16-byte chunk q (links, links, pos + weight, ray, dL/dcolour + dL/dsigma)
    __shared__ int4 ring[SAMPLE_THREADS / 8][QM_RING][5];
    int4 (*const my_ring)[5] = ring[threadIdx.x >> 3];
    const unsigned gmask = 0xffu << (threadIdx.x & 24);
    auto stage = [&](int s) {
        if (s < s1 && qi < 5)
            cp_async16(&my_ring[(s - s0) % QM_RING][qi],
                       reinterpret_cast<const int4*>(qv.rec + region_index<MODE>(rg, s, cur0)) + qi);
        cp_async_commit();   // (an empty group keeps the count per iteration uniform)
    };
#pragma unroll
    for (int k = 0; k < QM_RING; ++k) stage(s0 + k);

    for (int s = s0; s < s1; ++s) {
        cp_async_wait<QM_RING - 1>();   // this lane's chunk of record s has landed ...
        __syncwarp(gmask);              // ... and so have the other lanes'
        const int4(&slot)[5] = my_ring[(s - s0) % QM_RING];
        const float4 pw = *reinterpret_cast<const float4*>(&slot[2]);
        const int ray = slot[3].w;
        const float4 gg = *reinterpret_cast<const float4*>(&slot[4]);
        const int32_t ln = reinterpret_cast<const int32_t*>(&slot[0])[qi];   // corner qi of record s
        __syncwarp(gmask);              // everybody has read the slot: refill it
        stage(s + QM_RING);
        if (SLAB && pw.w < 0.f) continue;  // dead sample
        if (ray != cur_ray) {
            Y = __ldcg(qv.ray_Y + (int64_t)ray * 8 + qi);
            cur_ray = ray;
        }
        // same cell as the pending one?  Each lane compares its own corner and the group votes.
        const bool same = have && (__ballot_sync(gmask, ln == lown) & gmask) == gmask;
        if (!same) {
            if (have) flush();
            lown = ln;
            have = true;
            dacc = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
            // lane q pulls the grad row of corner q towards L2 for the flush one sample later
            if (SLOTS) {  // rank query for this lane's corner; the slot row is prefetched instead
                sown = ln >= 0 ? slot_of(sm, ln) : -1;
                if (sown >= 0) prefetch_l2(sm.grad + (size_t)(unsigned)sown * (unsigned)sm.width);
            } else if (ln >= 0) {
                const float* row = gr.grad_sh + (int64_t)ln * gsh_stride;
                prefetch_l2(row);
                prefetch_l2(row + (D - 1));
            }
        }
        float wq[8];
        corner_weights(pw.x, pw.y, pw.z, wq);
        const float gcs[3] = {gg.x, gg.y, gg.z};
        const float Ye[4] = {Y.x, Y.y, Y.z, Y.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gsel = ce[e] == 0 ? gcs[0] : (ce[e] == 1 ? gcs[1] : (ce[e] == 2 ? gcs[2] : 0.f));
            const float ge = Ye[e] * gsel;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q][e] = fmaf(wq[q], ge, acc[q][e]);
        }
        float wown = wq[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) wown = (qi == q) ? wq[q] : wown;
        dacc = fmaf(wown, gg.w, dacc);
    }
    if (have) flush();
}

// Resets the queue for a new batch.  The overflow flag of the PREVIOUS batch that used this
// workspace is folded into a sticky count first (header offset 192), so that callers that keep a
// workspace across steps and have no fallback (slab mode) can detect a dropped step at any later
// sync point.  (The single-GPU fused call re-does an overflowed batch with the lattice kernel.)
__global__ void clear_flags_kernel(int* counter, int* overflow, int* sticky) {
    if (*overflow) *sticky = (*sticky < 0 ? 0 : *sticky) + 1;
    *counter = 0;
    *overflow = 0;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

size_t queue_workspace_bytes(int64_t n_rays, int64_t records) {
    return 256 + align_up((size_t)n_rays * 4, 256) * 4 + align_up((size_t)n_rays * 8 * 16, 256) +
           align_up((size_t)n_rays * BLK_TABLE * 8, 256) + (size_t)records * sizeof(Record);
}

// Carves the workspace; false if it cannot hold even a minimal queue.
static bool carve_queue(void* workspace, size_t workspace_bytes, int64_t n_rays, QueueView& qv) {
    const size_t fixed = queue_workspace_bytes(n_rays, 0);
    if (!workspace || workspace_bytes < fixed + sizeof(Record) * 64 * (size_t)n_rays ||
        (reinterpret_cast<uintptr_t>(workspace) & 255u))
        return false;
    char* p = static_cast<char*>(workspace);
    qv.counter = reinterpret_cast<int*>(p);
    qv.overflow = reinterpret_cast<int*>(p + 128);
    qv.sticky = reinterpret_cast<int*>(p + 192);
    p += 256;
    qv.ray_first = reinterpret_cast<int*>(p);
    p += align_up((size_t)n_rays * 4, 256);
    qv.ray_logT = reinterpret_cast<float*>(p);
    p += align_up((size_t)n_rays * 4, 256);
    qv.ray_accum = reinterpret_cast<float*>(p);
    p += align_up((size_t)n_rays * 4, 256);
    qv.ray_Y = reinterpret_cast<float4*>(p);
    p += align_up((size_t)n_rays * 8 * 16, 256);
    qv.blk_table = reinterpret_cast<int2*>(p);
    p += align_up((size_t)n_rays * BLK_TABLE * 8, 256);
    qv.ray_nblk = reinterpret_cast<int*>(p);
    p += align_up((size_t)n_rays * 4, 256);
    qv.rec = reinterpret_cast<Record*>(p);
    const size_t cap = (workspace_bytes - fixed) / sizeof(Record);
    qv.capacity = (int)(cap > (size_t)0x7ff00000 ? (size_t)0x7ff00000 : cap);
    qv.counts = nullptr;
    qv.n_regions = 0;
    qv.region_cap = 0;
    return true;
}

// Returns false if the workspace is too small (caller uses the lattice path).
bool launch_queue_fused(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                        void* workspace, size_t workspace_bytes, const int** overflow_flag,
                        float** ray_logT, float** ray_accum, cudaStream_t st) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, a.n_rays, qv)) return false;
    *overflow_flag = qv.overflow;
    *ray_logT = qv.ray_logT;      // per-ray scratch the MSI fallback sequence may reuse
    *ray_accum = qv.ray_accum;
    const SlabView none{};

    clear_flags_kernel<<<1, 1, 0, st>>>(qv.counter, qv.overflow, qv.sticky);
    const unsigned ray_blocks = div_up(a.n_rays * WARP, TRACE_THREADS);
    if (a.counters) trace_kernel<true, 0><<<ray_blocks, TRACE_THREADS, 0, st>>>(g, opt, a, qv, none, ShipView{});
    else trace_kernel<false, 0><<<ray_blocks, TRACE_THREADS, 0, st>>>(g, opt, a, qv, none, ShipView{});
    const bool vec = march_vec_ok(g, &gr, a.gsh_stride);
    const unsigned sblocks = 148 * QM_SAMPLE_CTAS;  // persistent grid-stride CTAs of 256 threads
    if (vec) shade_kernel<true, 0, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, a.rgb_out, none, SlotMap{});
    else shade_kernel<false, 0, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, a.rgb_out, none, SlotMap{});
    // MSI scenes (cuvol.cu:1099-1149): the shells composite behind the foreground, so the background
    // forward pass sits between shade and seed (the seed needs the total colour) and its backward
    // pass takes over the accum the foreground recursion leaves.  Both skip an overflowed batch.
    const bool has_bg = g.bg_nlayers > 0;
    MarchArgs s = a;
    if (has_bg) {
        launch_bg_fwd(false, g, opt, a, qv.ray_logT, st, qv.overflow, 0);
        s.accum_out = qv.ray_accum;
    }
    seed_kernel<0><<<ray_blocks, TRACE_THREADS, 0, st>>>(opt, s, qv, none);
    if (vec) scatter_kernel<true, 0, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, gr, a.gsh_stride, a.gd_stride, SlotMap{});
    else scatter_kernel<false, 0, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, gr, a.gsh_stride, a.gd_stride, SlotMap{});
    if (has_bg && gr.grad_bg)
        launch_bg_bwd(g, opt, a, a.rgb_out, qv.ray_logT, qv.ray_accum, gr, st, qv.overflow, 0);
    return true;
}

// ---- fused train step (step.cu): the queue kernels with the compact gradient buffer ----------
// Phase A: trace + shade; the shade kernel also enters every touched row into sm.bits.  The caller
// then runs the TV marking pass and the rank scan (slot numbering) before phase B.
bool launch_step_trace_shade(const GridView& g, const Opts& opt, const MarchArgs& a, void* workspace,
                             size_t workspace_bytes, const SlotMap& sm, cudaStream_t st) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, a.n_rays, qv)) return false;
    if (!march_vec_ok(g, nullptr, 0)) return false;
    const SlabView none{};   // (the queue header was reset by the step's clear kernel)
    const unsigned ray_blocks = div_up(a.n_rays * WARP, TRACE_THREADS);
    if (a.counters) trace_kernel<true, 0><<<ray_blocks, TRACE_THREADS, 0, st>>>(g, opt, a, qv, none, ShipView{});
    else trace_kernel<false, 0><<<ray_blocks, TRACE_THREADS, 0, st>>>(g, opt, a, qv, none, ShipView{});
    shade_kernel<true, 0, true><<<148 * QM_SAMPLE_CTAS, SAMPLE_THREADS, 0, st>>>(g, qv, a.rgb_out, none, sm);
    return true;
}

// Queue header words, for the step's own clear kernel (same protocol as clear_flags_kernel).
bool step_queue_header(void* workspace, size_t workspace_bytes, int64_t n_rays, int** counter,
                       int** overflow, int** sticky) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, n_rays, qv)) return false;
    *counter = qv.counter;
    *overflow = qv.overflow;
    *sticky = qv.sticky;
    return true;
}

// Phase B: gradient seed per ray (needs the colours only) ...
bool launch_step_seed(const GridView& g, const Opts& opt, const MarchArgs& a, void* workspace,
                      size_t workspace_bytes, cudaStream_t st) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, a.n_rays, qv)) return false;
    const SlabView none{};
    seed_kernel<0><<<div_up(a.n_rays * WARP, TRACE_THREADS), TRACE_THREADS, 0, st>>>(opt, a, qv, none);
    return true;
}

// ... and, once the slots are numbered, the scatter into slot rows.
bool launch_step_scatter(const GridView& g, const MarchArgs& a, void* workspace, size_t workspace_bytes,
                         const SlotMap& sm, cudaStream_t st) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, a.n_rays, qv)) return false;
    scatter_kernel<true, 0, true><<<148 * QM_SAMPLE_CTAS, SAMPLE_THREADS, 0, st>>>(g, qv, GradView{}, 0, 0, sm);
    return true;
}

// ---- slab mode: the three phases between which the ranks exchange per-ray scalars -----------
bool launch_slab_trace(const GridView& g, const Opts& opt, const MarchArgs& a, void* workspace,
                       size_t workspace_bytes, int x_lo, int x_hi, float* logT_local,
                       cudaStream_t st) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, a.n_rays, qv)) return false;
    SlabView sl{};
    sl.x_lo = x_lo;
    sl.x_hi = x_hi;
    sl.logT_local = logT_local;
    clear_flags_kernel<<<1, 1, 0, st>>>(qv.counter, qv.overflow, qv.sticky);
    trace_kernel<false, 1><<<div_up(a.n_rays * WARP, TRACE_THREADS), TRACE_THREADS, 0, st>>>(
        g, opt, a, qv, sl, ShipView{});
    return true;
}

bool launch_slab_shade(const GridView& g, const Opts& opt, int64_t n_rays, void* workspace,
                       size_t workspace_bytes, const float* lt_in, float* part,
                       cudaStream_t st) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, n_rays, qv)) return false;
    SlabView sl{};
    sl.lt_in = lt_in;
    sl.stopped = part + 3;
    sl.stop_thresh = opt.stop_thresh;
    const unsigned sblocks = 148 * QM_SAMPLE_CTAS;
    if (march_vec_ok(g, nullptr, 0))
        shade_kernel<true, 1, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, part, sl, SlotMap{});
    else
        shade_kernel<false, 1, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, part, sl, SlotMap{});
    return true;
}

bool launch_slab_backward(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                          void* workspace, size_t workspace_bytes, const float* lt_in,
                          const float* gout, const float* accum_start, cudaStream_t st) {
    QueueView qv;
    if (!carve_queue(workspace, workspace_bytes, a.n_rays, qv)) return false;
    SlabView sl{};
    sl.lt_in = lt_in;
    sl.gout = gout;
    sl.accum_start = accum_start;
    sl.stop_thresh = opt.stop_thresh;
    seed_kernel<1><<<div_up(a.n_rays * WARP, TRACE_THREADS), TRACE_THREADS, 0, st>>>(opt, a, qv, sl);
    const unsigned sblocks = 148 * QM_SAMPLE_CTAS;
    if (march_vec_ok(g, &gr, a.gsh_stride))
        scatter_kernel<true, 1, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, gr, a.gsh_stride, a.gd_stride, SlotMap{});
    else
        scatter_kernel<false, 1, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, gr, a.gsh_stride, a.gd_stride, SlotMap{});
    return true;
}


// ---- ship mode: rays sharded for the traversal, records shipped to the slab that owns the cell ---
namespace {

// Per-rank symmetric workspace (identical layout on every rank; byte offsets):
//   0    queue header: [0] counter (unused), [128] overflow, [192] sticky overflow count
//   256  counts[16]: records rank k shipped to me this step (written by rank k)
//   384  cnt[16]: records I shipped to rank d so far (local)
//   512  ray_first[Qg] | ray_logT[Qg] | ray_Y[Qg][8] | rec[G][region_cap]
struct ShipWs {
    QueueView qv;
    int* counts;
    int* cnt;
};

bool carve_ship(void* workspace, size_t bytes, int64_t n_rays_global, int n_slabs, int64_t region_cap,
                ShipWs& w) {
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) || n_slabs < 1 ||
        n_slabs > SVOX2_MAX_PEERS || region_cap < 32 ||
        bytes < ship_workspace_bytes(n_rays_global, n_slabs, region_cap))
        return false;
    char* p = static_cast<char*>(workspace);
    QueueView& qv = w.qv;
    qv.counter = reinterpret_cast<int*>(p);
    qv.overflow = reinterpret_cast<int*>(p + 128);
    qv.sticky = reinterpret_cast<int*>(p + 192);
    w.counts = reinterpret_cast<int*>(p + 256);
    w.cnt = reinterpret_cast<int*>(p + 384);
    p += 512;
    qv.ray_first = reinterpret_cast<int*>(p);
    p += align_up((size_t)n_rays_global * 4, 256);
    qv.ray_logT = reinterpret_cast<float*>(p);
    p += align_up((size_t)n_rays_global * 4, 256);
    qv.ray_accum = nullptr;
    qv.blk_table = nullptr;
    qv.ray_nblk = nullptr;
    qv.ray_Y = reinterpret_cast<float4*>(p);
    p += align_up((size_t)n_rays_global * 8 * 16, 256);
    qv.rec = reinterpret_cast<Record*>(p);
    qv.capacity = (int)(region_cap * n_slabs);
    qv.counts = w.counts;
    qv.n_regions = n_slabs;
    qv.region_cap = (int)region_cap;
    return true;
}

// SH basis table of every ray of the global batch (the sample-parallel kernels of this rank may see
// samples of any ray)
__global__ void __launch_bounds__(256)
ray_basis_kernel(GridView g, const float* __restrict__ dirs, int64_t n_rays, float4* __restrict__ ray_Y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ray = t >> 3;
    const int q = (int)(t & 7);
    if (ray >= n_rays) return;
    float dw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) dw[i] = __ldg(dirs + ray * 3 + i);
    float y[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = q * 4 + e;
        y[e] = j < g.sh_dim ? sh_basis_k(g.basis_dim, j % g.basis_dim, dw) : 0.f;
    }
    ray_Y[ray * 8 + q] = make_float4(y[0], y[1], y[2], y[3]);
}

struct PeerInts {
    int* p[SVOX2_MAX_PEERS];
};
// rank `me` tells every peer how many records it shipped there
__global__ void ship_publish_kernel(PeerInts counts, const int* __restrict__ cnt, int n, int me, int cap) {
    const int d = threadIdx.x;
    if (d < n) counts.p[d][me] = min(cnt[d], cap);
}

// end of a step: forget this step's records (overflow folded into the sticky count first)
__global__ void ship_reset_kernel(int* overflow, int* sticky, int* counts, int* cnt) {
    const int t = threadIdx.x;
    if (t == 0) {
        if (*overflow) *sticky = (*sticky < 0 ? 0 : *sticky) + 1;
        *overflow = 0;
    }
    if (t < SVOX2_MAX_PEERS) {
        counts[t] = 0;
        cnt[t] = 0;
    }
}

}  // namespace

size_t ship_workspace_bytes(int64_t n_rays_global, int n_slabs, int64_t region_cap) {
    return 512 + align_up((size_t)n_rays_global * 4, 256) * 2 + align_up((size_t)n_rays_global * 8 * 16, 256) +
           (size_t)n_slabs * (size_t)region_cap * sizeof(Record);
}

bool launch_ship_trace(const GridView& g, const Opts& opt, const MarchArgs& a, int64_t ray_offset,
                       int64_t n_rays_global, const int* bounds, int n_slabs, int rank,
                       const uint64_t* peer_ws, size_t workspace_bytes, int64_t region_cap, cudaStream_t st) {
    ShipWs mine;
    if (!carve_ship(reinterpret_cast<void*>(peer_ws[rank]), workspace_bytes, n_rays_global, n_slabs, region_cap, mine))
        return false;
    ShipView sv{};
    PeerInts pc{};
    for (int d = 0; d < n_slabs; ++d) {
        ShipWs w;
        if (!carve_ship(reinterpret_cast<void*>(peer_ws[d]), workspace_bytes, n_rays_global, n_slabs, region_cap, w))
            return false;
        sv.rec[d] = w.qv.rec + (size_t)rank * (size_t)region_cap;
        sv.ray_first[d] = w.qv.ray_first;
        sv.ray_logT[d] = w.qv.ray_logT;
        sv.overflow[d] = w.qv.overflow;
        pc.p[d] = w.counts;
    }
    for (int d = 0; d <= n_slabs; ++d) sv.bounds[d] = bounds[d];
    sv.cnt = mine.cnt;
    sv.n = n_slabs;
    sv.region_base = rank * (int)region_cap;
    sv.region_cap = (int)region_cap;
    sv.ray_offset = ray_offset;
    ray_basis_kernel<<<div_up(n_rays_global * 8, 256), 256, 0, st>>>(g, a.dirs, n_rays_global, mine.qv.ray_Y);
    const SlabView none{};
    const unsigned ray_blocks = div_up(a.n_rays * WARP, TRACE_THREADS);
    if (a.counters) trace_kernel<true, 2><<<ray_blocks, TRACE_THREADS, 0, st>>>(g, opt, a, mine.qv, none, sv);
    else trace_kernel<false, 2><<<ray_blocks, TRACE_THREADS, 0, st>>>(g, opt, a, mine.qv, none, sv);
    ship_publish_kernel<<<1, 32, 0, st>>>(pc, mine.cnt, n_slabs, rank, (int)region_cap);
    return true;
}

bool launch_ship_shade(const GridView& g, int64_t n_rays_global, void* workspace, size_t workspace_bytes,
                       int n_slabs, int64_t region_cap, float* part, cudaStream_t st) {
    ShipWs w;
    if (!carve_ship(workspace, workspace_bytes, n_rays_global, n_slabs, region_cap, w)) return false;
    const SlabView none{};
    const unsigned sblocks = 148 * QM_SAMPLE_CTAS;
    if (march_vec_ok(g, nullptr, 0))
        shade_kernel<true, 2, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, w.qv, part, none, SlotMap{});
    else
        shade_kernel<false, 2, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, w.qv, part, none, SlotMap{});
    return true;
}

const float* ship_ray_logT(void* workspace, size_t workspace_bytes, int64_t n_rays_global, int n_slabs,
                           int64_t region_cap) {
    ShipWs w;
    if (!carve_ship(workspace, workspace_bytes, n_rays_global, n_slabs, region_cap, w)) return nullptr;
    return w.qv.ray_logT;
}

bool launch_ship_backward(const GridView& g, const Opts& opt, const MarchArgs& a, const GradView& gr,
                          void* workspace, size_t workspace_bytes, int n_slabs, int64_t region_cap,
                          const float* gout, const float* accum_start, cudaStream_t st) {
    ShipWs w;
    if (!carve_ship(workspace, workspace_bytes, a.n_rays, n_slabs, region_cap, w)) return false;
    SlabView sl{};
    sl.gout = gout;
    sl.accum_start = accum_start;
    seed_kernel<2><<<div_up(a.n_rays * WARP, TRACE_THREADS), TRACE_THREADS, 0, st>>>(opt, a, w.qv, sl);
    const unsigned sblocks = 148 * QM_SAMPLE_CTAS;
    if (march_vec_ok(g, &gr, a.gsh_stride))
        scatter_kernel<true, 2, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, w.qv, gr, a.gsh_stride, a.gd_stride, SlotMap{});
    else
        scatter_kernel<false, 2, false><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, w.qv, gr, a.gsh_stride, a.gd_stride, SlotMap{});
    ship_reset_kernel<<<1, 32, 0, st>>>(w.qv.overflow, w.qv.sticky, w.counts, w.cnt);
    return true;
}


// ---- multi-GPU step on the compact gradient buffer (step.cu): the sample-parallel kernels of the
// ---- slab (MODE 1) / ship (MODE 2) step with slot marking and slot scatter ----------------------
static bool resolve_queue(const QueueRef& q, QueueView& qv) {
    if (q.n_slabs > 0) {
        ShipWs w;
        if (!carve_ship(q.ws, q.bytes, q.n_rays, q.n_slabs, q.region_cap, w)) return false;
        qv = w.qv;
        return true;
    }
    return carve_queue(q.ws, q.bytes, q.n_rays, qv);
}

bool launch_mg_shade(const GridView& g, const Opts& opt, const QueueRef& q, const float* lt_in, float* part,
                     const SlotMap& sm, cudaStream_t st) {
    QueueView qv;
    if (!resolve_queue(q, qv) || !march_vec_ok(g, nullptr, 0)) return false;
    SlabView sl{};
    sl.lt_in = lt_in;
    sl.stopped = part + 3;
    sl.stop_thresh = opt.stop_thresh;
    const unsigned sblocks = 148 * QM_SAMPLE_CTAS;
    if (q.n_slabs > 0) shade_kernel<true, 2, true><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, part, sl, sm);
    else shade_kernel<true, 1, true><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, part, sl, sm);
    return true;
}

bool launch_mg_seed(const Opts& opt, const MarchArgs& a, const QueueRef& q, const float* lt_in,
                    const float* gout, const float* accum_start, cudaStream_t st) {
    QueueView qv;
    if (!resolve_queue(q, qv)) return false;
    SlabView sl{};
    sl.lt_in = lt_in;
    sl.gout = gout;
    sl.accum_start = accum_start;
    sl.stop_thresh = opt.stop_thresh;
    const unsigned blocks = div_up(a.n_rays * WARP, TRACE_THREADS);
    if (q.n_slabs > 0) seed_kernel<2><<<blocks, TRACE_THREADS, 0, st>>>(opt, a, qv, sl);
    else seed_kernel<1><<<blocks, TRACE_THREADS, 0, st>>>(opt, a, qv, sl);
    return true;
}

bool launch_mg_scatter(const GridView& g, const QueueRef& q, const SlotMap& sm, cudaStream_t st) {
    QueueView qv;
    if (!resolve_queue(q, qv)) return false;
    const unsigned sblocks = 148 * QM_SAMPLE_CTAS;
    if (q.n_slabs > 0) {
        scatter_kernel<true, 2, true><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, GradView{}, 0, 0, sm);
        ShipWs w;
        carve_ship(q.ws, q.bytes, q.n_rays, q.n_slabs, q.region_cap, w);
        ship_reset_kernel<<<1, 32, 0, st>>>(w.qv.overflow, w.qv.sticky, w.counts, w.cnt);
    } else {
        scatter_kernel<true, 1, true><<<sblocks, SAMPLE_THREADS, 0, st>>>(g, qv, GradView{}, 0, 0, sm);
    }
    return true;
}

}  // namespace svox2b
