// The reference's ray traversal (render_lerp_kernel_cuvol.cu:56-109) in warp-cooperative form,
// shared by the lattice kernel (march.cu) and the sample-queue trace kernel (queue_march.cu).
//
// One ROUND looks at 32 lattice candidates: lane k holds t_0 advanced by k sequentially rounded
// adds (bit-identical to the reference's `t += step` chain), locates its cell and loads its corner
// links.  Ballots then decide which candidates are on the reference's sample lattice: a skip at
// lane j jumps to t_j + ceil(skip/step)*step; if that value is bit-equal to a later candidate the
// round continues there, otherwise the next round starts at it.  Member lanes interpolate sigma;
// the running log-transmittance is the only sequential part.
#pragma once
#include "common.cuh"

namespace svox2b {

__device__ __forceinline__ unsigned range_mask(int a, int b) {  // bits a .. b-1, 0 <= a <= b <= 32
    const unsigned hi = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
    const unsigned lo = (a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
    return hi & ~lo;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// t0 advanced by `lane` sequentially rounded adds of `step` -- the reference's `t += step` chain.
// Inside one binade every add moves the mantissa by the same integer number D of ulps (round to
// nearest of step = S ulp + r: D = S for r < ulp/2, S + 1 for r > ulp/2), so the chain is the exact
// arithmetic progression t0 + k*(t1 - t0): t1 - t0 is exact (Sterbenz) and so is the fma as long as
// t_31 stays in t0's binade.  The two cases where that fails -- the progression leaves the binade
// (ulp changes), or r == ulp/2 exactly (ties round to even, so the first add may differ from the
// later ones) -- are detected warp-uniformly and take the 31 dependent adds.
#ifndef TRAVERSE_FAST_LATTICE
#define TRAVERSE_FAST_LATTICE 1
#endif
__device__ __forceinline__ float lattice_position(float t0, float step, int lane) {
#if TRAVERSE_FAST_LATTICE
    const float t1 = __fadd_rn(t0, step);
    const float delta = __fsub_rn(t1, t0);
    const float t31 = __fmaf_rn(31.f, delta, t0);
    const int e0 = __float_as_int(t0) & 0x7f800000;                 // sign bit set (t0 < 0) fails the test below
    const bool in_binade = __float_as_int(t0) == (e0 | (__float_as_int(t0) & 0x007fffff)) &&
                           (__float_as_int(t31) & 0xff800000) == e0 && e0 >= (64 << 23) && e0 <= (190 << 23);
    const float half_ulp = __int_as_float(e0 - (24 << 23));
    const bool tie = fabsf(__fsub_rn(step, delta)) == half_ulp;
    if (in_binade && !tie && delta > 0.f) return __fmaf_rn((float)lane, delta, t0);
#endif
    float t = t0;
#pragma unroll
    for (int i = 0; i < 31; ++i) {
        const float nx = __fadd_rn(t, step);
        t = (i < lane) ? nx : t;
    }
    return t;
}

struct Candidates {
    float tk;        // this lane's lattice position
    bool valid;      // tk <= tmax
    int l[3];        // cell
    float pos[3];    // position inside the cell
    int32_t lk[8];   // corner links (-1 where not loaded)
    float skipd;     // empty-space skip distance of the cell (0: none)
    bool is_skip;
};

// TWO_PHASE: load the skip-deciding link first and the other seven only for candidates that do
// not skip (fewer link sectors in empty space, one more dependent round trip for samples).
// x_lo / x_hi (TWO_PHASE only): the other seven links are loaded only for candidates whose cell lies
// in [x_lo, x_hi) -- slab mode traverses the part of a ray outside its slab for the lattice alone.
template <bool TWO_PHASE>
__device__ __forceinline__ void load_candidates(const Ray& ray, const GridView& g, float t0,
                                                float step, int lane, Candidates& c,
                                                int x_lo = -0x7fffffff, int x_hi = 0x7fffffff) {
    const float t = lattice_position(t0, step, lane);
    c.tk = t;
    c.valid = t <= ray.tmax;
    locate(ray, g, t, c.l, c.pos);
    const int32_t* lp = g.links + ((int64_t)c.l[0] * g.offx + (int64_t)c.l[1] * g.offy + c.l[2]);
#pragma unroll
    for (int q = 0; q < 8; ++q) c.lk[q] = -1;
    if (TWO_PHASE) {
        if (c.valid) c.lk[0] = __ldg(lp);
        c.skipd = skip_dist(ray, c.valid ? c.lk[0] : 0, c.l, c.pos);
        c.is_skip = c.valid && (c.skipd >= step);
        if (c.valid && !c.is_skip && c.l[0] >= x_lo && c.l[0] < x_hi) {
#pragma unroll
            for (int q = 1; q < 8; ++q) c.lk[q] = __ldg(lp + corner_off(g, q));
        }
    } else {
        if (c.valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q) c.lk[q] = __ldg(lp + corner_off(g, q));
        }
        c.skipd = skip_dist(ray, c.valid ? c.lk[0] : 0, c.l, c.pos);
        c.is_skip = c.valid && (c.skipd >= step);
    }
}

// Which candidates are samples of the reference march (warp-uniform result).  Sets t_next (start
// of the next round) or done (ray left [tmin, tmax]); n_skips counts the skips taken.
__device__ __forceinline__ unsigned resolve_lattice(const Ray& ray, const Candidates& c, float step,
                                                    float& t_next, bool& done, unsigned& n_skips) {
    const unsigned valid_m = __ballot_sync(FULL, c.valid);
    const unsigned skip_m = __ballot_sync(FULL, c.is_skip);
    unsigned members = 0;
    int cur = 0;
    for (;;) {
        const unsigned rem = 0xffffffffu << cur;
        const unsigned inv = ~valid_m & rem, sk = skip_m & rem;
        const int first_inv = inv ? __ffs(inv) - 1 : 32;
        const int first_sk = sk ? __ffs(sk) - 1 : 32;
        if (first_inv <= first_sk) {  // ray leaves [tmin,tmax] (or the round is exhausted) first
            members |= range_mask(cur, first_inv);
            if (first_inv < 32) done = true;
            else t_next = __fadd_rn(__shfl_sync(FULL, c.tk, 31), step);
            break;
        }
        const int j = first_sk;  // lane j skips empty space
        members |= range_mask(cur, j);
        ++n_skips;
        const float tn = skip_advance(__shfl_sync(FULL, c.tk, j), __shfl_sync(FULL, c.skipd, j), step);
        const unsigned later = (j < 31) ? (0xffffffffu << (j + 1)) : 0u;
        const unsigned eq = __ballot_sync(FULL, c.tk == tn) & later;
        if (eq) {  // landed exactly on a later candidate of this round: keep going there
            cur = __ffs(eq) - 1;
            continue;
        }
        t_next = tn;
        if (!(tn <= ray.tmax)) done = true;
        break;
    }
    return members;
}

// Touch the link lines the next round's candidates will probe (positions need not be exact).
__device__ __forceinline__ void prefetch_next_links(const Ray& ray, const GridView& g, float t_next,
                                                    float step, int lane, int x_lo = -0x7fffffff,
                                                    int x_hi = 0x7fffffff) {
    const float ta = fmaf((float)lane, step, t_next);
    if (ta <= ray.tmax) {
        int nl[3];
        float npos[3];
        locate(ray, g, ta, nl, npos);
        const int32_t* np = g.links + ((int64_t)nl[0] * g.offx + (int64_t)nl[1] * g.offy + nl[2]);
        prefetch_l2(np);
        if (nl[0] >= x_lo - 1 && nl[0] <= x_hi) {   // outside the slab only the skip-deciding link is read
            prefetch_l2(np + g.offy);
            prefetch_l2(np + g.offx);
            prefetch_l2(np + g.offx + g.offy);
        }
    }
}

__device__ __forceinline__ float member_sigma(const GridView& g, const Candidates& c, bool member) {
    float sigma = 0.f;
    if (member) {
        float sv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            sv[q] = c.lk[q] >= 0 ? __ldg(g.density + (int64_t)c.lk[q] * g.density_stride) : 0.f;
        sigma = trilerp8(sv, c.pos);
    }
    return sigma;
}

// Running log-transmittance over the round's active samples in sample order.  Only the subtraction
// chain is sequential (bit-exact running sum); the exp / stop test (cuvol.cu:103-106, :367-369) of
// all samples runs in parallel afterwards.  Returns the active mask, truncated behind the sample
// that fired the stop threshold (members likewise); my_before = log-transmittance before this
// lane's sample.
__device__ __forceinline__ unsigned transmittance_chain(bool active, float pcnt, float stop_thresh,
                                                        int lane, float& lt, unsigned& members,
                                                        bool& done, bool& stopped, float& my_before) {
    unsigned act_m = __ballot_sync(FULL, active);
    my_before = 0.f;
    unsigned it = act_m;
    while (it) {
        const int i = __ffs(it) - 1;
        it &= it - 1;
        const float p = __shfl_sync(FULL, pcnt, i);
        my_before = (lane == i) ? lt : my_before;
        lt = __fsub_rn(lt, p);
    }
    const float aft = __fsub_rn(my_before, pcnt);
    const unsigned tm = __ballot_sync(FULL, ((act_m >> lane) & 1u) && fexp(aft) < stop_thresh);
    if (tm) {
        const int sidx = __ffs(tm) - 1;
        const unsigned upto = (2u << sidx) - 1u;
        act_m &= upto;
        members &= upto;
        done = true;
        stopped = true;
        lt = __shfl_sync(FULL, aft, sidx);
    }
    return act_m;
}

}  // namespace svox2b
