/*
 * svox2_oracle.c -- CPU restatement of the sxyu/svox2 hot path.  TEST INFRASTRUCTURE ONLY.
 * See svox2_oracle.h for scope, pinning and known deviations.
 *
 * Compile with -ffp-contract=off: every fused multiply-add that the reference's device
 * code performs is written as an explicit fmaf() here (positions, lerps, transform_coord);
 * everything else is separately rounded, which is what the reference's sm_100 SASS does on
 * the parity-critical chain (t update, pcnt, log_transmit: FMUL/FADD, checked with
 * cuobjdump on the reference build).
 */
#include "svox2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ---------------------------------------------------------------------------------------
 * small helpers
 * ------------------------------------------------------------------------------------- */

/* cuda_util.cuh:72-75 : lerp(a,b,w) = fmaf(w, b-a, a) */
static inline float lerpf(float a, float b, float w) { return fmaf(w, b - a, a); }

/* Multi-threading for the bench's cpu_baseline leg: the ray loops run on g_threads pthreads
 * (dynamic chunks of 16 rays); gradient accumulation then uses a CAS loop.  With one thread
 * (the default, what the tests use) everything is plain sequential code. */
static int g_threads = 1;
static inline void add_double(double* p, double v) {
    if (g_threads <= 1) { *p += v; return; }
    uint64_t* pu = (uint64_t*)p;
    uint64_t old = __atomic_load_n(pu, __ATOMIC_RELAXED), neu;
    do {
        double d;
        memcpy(&d, &old, 8);
        d += v;
        memcpy(&neu, &d, 8);
    } while (!__atomic_compare_exchange_n(pu, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
typedef void (*range_fn)(int64_t begin, int64_t end, void* ctx);
typedef struct { range_fn fn; void* ctx; int64_t n; int64_t next; } pfor_t;
static void* pfor_worker(void* arg) {
    pfor_t* p = (pfor_t*)arg;
    for (;;) {
        const int64_t b = __atomic_fetch_add(&p->next, 16, __ATOMIC_RELAXED);
        if (b >= p->n) break;
        p->fn(b, b + 16 < p->n ? b + 16 : p->n, p->ctx);
    }
    return NULL;
}
static void parallel_for(int64_t n, range_fn fn, void* ctx) {
    if (g_threads <= 1 || n <= 16) { fn(0, n, ctx); return; }
    pfor_t p = {fn, ctx, n, 0};
    pthread_t th[256];
    const int nt = g_threads > 256 ? 256 : g_threads;
    for (int i = 1; i < nt; ++i) pthread_create(&th[i], NULL, pfor_worker, &p);
    pfor_worker(&p);
    for (int i = 1; i < nt; ++i) pthread_join(th[i], NULL);
}
static inline float minf(float a, float b) { return a < b ? a : b; }
static inline float maxf(float a, float b) { return a > b ? a : b; }
static inline int mini(int a, int b) { return a < b ? a : b; }
static inline int maxi(int a, int b) { return a > b ? a : b; }

/* rnorm3df / norm3df (cuda_util.cuh:77-87): evaluated in double, rounded once. */
static inline float rnorm3(const float* v) {
    double s = (double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2];
    return (float)(1.0 / sqrt(s));
}
static inline float norm3(const float* v) {
    double s = (double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2];
    return (float)sqrt(s);
}
/* _EXP == __expf on device (util.hpp:14-17); expf here. */
static inline float EXPF(float x) { return expf(x); }

static inline float dot3(const float* a, const float* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* SH constants, render_util.cuh:300-318 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};

/* calc_sh, render_util.cuh:332-363 */
static void calc_sh(int basis_dim, const float* dir, float* out) {
    out[0] = SH_C0;
    const float x = dir[0], y = dir[1], z = dir[2];
    const float xx = x * x, yy = y * y, zz = z * z;
    const float xy = x * y, yz = y * z, xz = x * z;
    if (basis_dim >= 9) {
        out[4] = SH_C2[0] * xy;
        out[5] = SH_C2[1] * yz;
        out[6] = (float)((double)SH_C2[2] * (2.0 * (double)zz - (double)xx - (double)yy));
        out[7] = SH_C2[3] * xz;
        out[8] = SH_C2[4] * (xx - yy);
    }
    if (basis_dim >= 4) {
        out[1] = -SH_C1 * y;
        out[2] = SH_C1 * z;
        out[3] = -SH_C1 * x;
    }
}

/* ConcentricSpheresIntersector, render_util.cuh:546-576 */
typedef struct { float q2a, qb, f; } csi_t;
static csi_t csi_make(const float* origin, const float* dir) {
    csi_t c;
    c.q2a = 2 * dot3(dir, dir);
    c.qb = 2 * dot3(origin, dir);
    c.f = c.qb * c.qb - 2 * c.q2a * dot3(origin, origin);
    return c;
}
static int csi_intersect(const csi_t* c, float r, float* out, int near) {
    float det = c->f + 2 * c->q2a * r * r;
    if (det < 0) return 0;
    if (near) *out = (-c->qb - sqrtf(det)) / c->q2a;
    else *out = (-c->qb + sqrtf(det)) / c->q2a;
    return 1;
}

typedef struct {
    float origin[3], dir[3];
    float tmin, tmax, world_step;
    float pos[3];
    int32_t l[3];
} ray_t;

/* ray_find_bounds, render_util.cuh:578-628 (+ transform_coord cuda_util.cuh:61-67,
 * _get_delta_scale render_util.cuh:463-474) */
static void ray_find_bounds(ray_t* ray, const og_grid_t* g, const og_opt_t* opt) {
    for (int i = 0; i < 3; ++i) ray->origin[i] = fmaf(ray->origin[i], g->scaling[i], g->offset[i]);
    for (int i = 0; i < 3; ++i) ray->dir[i] *= g->scaling[i];
    const float delta_scale = rnorm3(ray->dir);
    for (int i = 0; i < 3; ++i) ray->dir[i] *= delta_scale;
    ray->world_step = delta_scale * opt->step_size;

    if (opt->use_spheric_clip) {
        float so[3], sd[3];
        for (int i = 0; i < 3; ++i) {
            const float ss = 2.f / (float)g->size[i];
            so[i] = fmaf(ray->origin[i] + 0.5f, ss, -1.f);
            sd[i] = ray->dir[i] * ss;
        }
        csi_t c = csi_make(so, sd);
        if (!csi_intersect(&c, 1.f, &ray->tmax, 0) ||
            !csi_intersect(&c, 1.f - opt->near_clip, &ray->tmin, 1)) {
            ray->tmin = 1e-9f;
            ray->tmax = 0.f;
        }
    } else {
        ray->tmin = opt->near_clip / ray->world_step * opt->step_size;
        ray->tmax = 2e3f;
        for (int i = 0; i < 3; ++i) {
            const float invdir = 1.0f / ray->dir[i]; /* (float)(1.0/d) == IEEE f32 reciprocal */
            const float t1 = (-0.5f - ray->origin[i]) * invdir;
            const float t2 = ((float)g->size[i] - 0.5f - ray->origin[i]) * invdir;
            if (ray->dir[i] != 0.f) {
                ray->tmin = maxf(ray->tmin, minf(t1, t2));
                ray->tmax = minf(ray->tmax, maxf(t1, t2));
            }
        }
    }
}

/* cuvol.cu:57-63 : position, clamp, cell index, fractional part */
static inline void ray_locate(ray_t* ray, const og_grid_t* g, float t) {
    for (int j = 0; j < 3; ++j) {
        float p = fmaf(t, ray->dir[j], ray->origin[j]);
        p = minf(maxf(p, 0.f), (float)g->size[j] - 1.f);
        ray->l[j] = mini((int32_t)p, g->size[j] - 2);
        ray->pos[j] = p - (float)ray->l[j];
    }
}

/* compute_skip_dist, render_util.cuh:235-295 */
static float compute_skip_dist(const ray_t* ray, const og_grid_t* g) {
    const int64_t offx = (int64_t)g->size[1] * g->size[2], offy = g->size[2];
    const int32_t link_val = g->links[offx * ray->l[0] + offy * ray->l[1] + ray->l[2]];
    if (link_val >= -1) return 0.f;
    const uint32_t dist = (uint32_t)(-link_val);
    const uint32_t shift = dist - 1;
    const uint32_t side = (uint32_t)((float)(1 << shift) - 1.f);
    float tmin = 0.f, tmax = 1e9f;
    for (int i = 0; i < 3; ++i) {
        int ul = ((ray->l[i] >> shift) << shift);
        ul -= ray->l[i];
        const float invdir = 1.0f / ray->dir[i];
        const float t1 = ((float)ul - ray->pos[i]) * invdir;
        const float t2 = ((float)((uint32_t)ul + side) - ray->pos[i]) * invdir;
        if (ray->dir[i] != 0.f) {
            tmin = maxf(tmin, minf(t1, t2));
            tmax = minf(tmax, maxf(t1, t2));
        }
    }
    if (tmin > 0.f) return 0.f;
    return tmax;
}

/* 8 corner links of the current cell, in the reference's visiting order
 * (0, 1, offy, offy+1, offx, offx+1, offx+offy, offx+offy+1). */
static inline void corner_links(const ray_t* ray, const og_grid_t* g, int32_t* lk) {
    const int64_t offx = (int64_t)g->size[1] * g->size[2], offy = g->size[2];
    const int32_t* p = g->links + (offx * ray->l[0] + offy * ray->l[1] + ray->l[2]);
    lk[0] = p[0];          lk[1] = p[1];
    lk[2] = p[offy];       lk[3] = p[offy + 1];
    lk[4] = p[offx];       lk[5] = p[offx + 1];
    lk[6] = p[offx + offy]; lk[7] = p[offx + offy + 1];
}

/* trilerp_cuvol_one, render_util.cuh:72-91 : z, then y, then x */
static inline float trilerp8(const float* v, const float* pos) {
    const float ix0y0 = lerpf(v[0], v[1], pos[2]);
    const float ix0y1 = lerpf(v[2], v[3], pos[2]);
    const float ix0 = lerpf(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = lerpf(v[4], v[5], pos[2]);
    const float ix1y1 = lerpf(v[6], v[7], pos[2]);
    const float ix1 = lerpf(ix1y0, ix1y1, pos[1]);
    return lerpf(ix0, ix1, pos[0]);
}
static inline float trilerp_links(const int32_t* lk, const float* data, int64_t stride, int idx,
                                  const float* pos) {
    float v[8];
    for (int c = 0; c < 8; ++c) v[c] = lk[c] >= 0 ? data[(int64_t)lk[c] * stride + idx] : 0.f;
    return trilerp8(v, pos);
}
/* trilerp_backward_cuvol_one weights, render_util.cuh:94-121 */
static inline void trilerp_weights(const float* pos, float g, float* w) {
    const float ay = 1.f - pos[1], az = 1.f - pos[2];
    float xo = (1.0f - pos[0]) * g;
    w[0] = ay * az * xo;  w[1] = ay * pos[2] * xo;
    w[2] = pos[1] * az * xo;  w[3] = pos[1] * pos[2] * xo;
    xo = pos[0] * g;
    w[4] = ay * az * xo;  w[5] = ay * pos[2] * xo;
    w[6] = pos[1] * az * xo;  w[7] = pos[1] * pos[2] * xo;
}

/* ---------------------------------------------------------------------------------------
 * foreground march, forward: trace_ray_cuvol, cuvol.cu:30-120
 * ------------------------------------------------------------------------------------- */
static void trace_ray_fwd(const og_grid_t* g, const og_opt_t* opt, ray_t* ray, const float* sph,
                          float* out, float* out_logT, og_counts_t* cnt) {
    const int B = g->basis_dim, D = 3 * B;
    const int has_bg = g->bg_nlayers > 0;
    if (ray->tmin > ray->tmax) {
        for (int c = 0; c < 3; ++c) out[c] = has_bg ? 0.f : opt->background_brightness;
        if (out_logT) *out_logT = 0.f;
        return;
    }
    float t = ray->tmin;
    float outv[3] = {0.f, 0.f, 0.f};
    float log_transmit = 0.f;
    while (t <= ray->tmax) {
        ray_locate(ray, g, t);
        const float skip = compute_skip_dist(ray, g);
        if (skip >= opt->step_size) {
            t += ceilf(skip / opt->step_size) * opt->step_size;
            if (cnt) cnt->s_skip++;
            continue;
        }
        int32_t lk[8];
        corner_links(ray, g, lk);
        const float sigma = trilerp_links(lk, g->density, 1, 0, ray->pos);
        if (cnt) cnt->s_vis++;
        if (opt->last_sample_opaque && t + opt->step_size > ray->tmax) ray->world_step = 1e9f;
        if (sigma > opt->sigma_thresh) {
            if (cnt) cnt->s_act++;
            float col[3] = {0.f, 0.f, 0.f};
            for (int j = 0; j < D; ++j) {
                float lc = trilerp_links(lk, g->sh, D, j, ray->pos);
                col[j / B] += lc * sph[j % B];
            }
            const float pcnt = ray->world_step * sigma;
            const float weight = EXPF(log_transmit) * (1.f - EXPF(-pcnt));
            log_transmit -= pcnt;
            for (int c = 0; c < 3; ++c) outv[c] = fmaf(weight, maxf(col[c] + 0.5f, 0.f), outv[c]);
            if (EXPF(log_transmit) < opt->stop_thresh) {
                log_transmit = -1e3f;
                break;
            }
        }
        t += opt->step_size;
    }
    if (!has_bg) {
        const float T = EXPF(log_transmit);
        for (int c = 0; c < 3; ++c) outv[c] = fmaf(T, opt->background_brightness, outv[c]);
    }
    if (out_logT) *out_logT = log_transmit;
    for (int c = 0; c < 3; ++c) out[c] = outv[c];
}

/* ---------------------------------------------------------------------------------------
 * foreground march, backward: trace_ray_cuvol_backward, cuvol.cu:227-383
 * gradients accumulate into double buffers.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    double* gd;   /* [N] */
    double* gsh;  /* [N*D] */
    double* gbg;  /* [Nbg*L*4] or NULL */
    float* fgd;   /* float-accumulation mode (bench): accumulate straight into the outputs */
    float* fgsh;
    float* fgbg;
    uint8_t* mask;
    uint8_t* mask_bg;
} grads_t;

static int g_accum_float = 0;
static inline void add_float(float* p, float v) {
    if (g_threads <= 1) { *p += v; return; }
    uint32_t* pu = (uint32_t*)p;
    uint32_t old = __atomic_load_n(pu, __ATOMIC_RELAXED), neu;
    do {
        float d;
        memcpy(&d, &old, 4);
        d += v;
        memcpy(&neu, &d, 4);
    } while (!__atomic_compare_exchange_n(pu, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
static inline void acc_sh(grads_t* gr, int64_t i, float v) {
    if (gr->fgsh) add_float(&gr->fgsh[i], v); else add_double(&gr->gsh[i], (double)v);
}
static inline void acc_d(grads_t* gr, int64_t i, float v) {
    if (gr->fgd) add_float(&gr->fgd[i], v); else add_double(&gr->gd[i], (double)v);
}
static inline void acc_bg(grads_t* gr, int64_t i, float v) {
    if (gr->fgbg) add_float(&gr->fgbg[i], v); else if (gr->gbg) add_double(&gr->gbg[i], (double)v);
}

static void trace_ray_bwd(const og_grid_t* g, const og_opt_t* opt, ray_t* ray, const float* sph,
                          const float* grad_output, const float* color_cache,
                          float log_transmit_in, float beta_loss, float sparsity_loss,
                          grads_t* gr, float* accum_out, float* logT_out) {
    const int B = g->basis_dim, D = 3 * B;
    float accum = fmaf(color_cache[0], grad_output[0],
                       fmaf(color_cache[1], grad_output[1], color_cache[2] * grad_output[2]));
    if (ray->tmin > ray->tmax) {
        if (accum_out) *accum_out = accum;
        if (logT_out) *logT_out = 0.f;
        return;
    }
    if (beta_loss > 0.f) {
        const float transmit_in = EXPF(log_transmit_in);
        beta_loss = (float)((double)beta_loss *
                            (1 - (double)transmit_in / (1 - (double)transmit_in + 1e-3)));
        accum += beta_loss;
    }
    float t = ray->tmin;
    float log_transmit = 0.f;
    while (t <= ray->tmax) {
        ray_locate(ray, g, t);
        const float skip = compute_skip_dist(ray, g);
        if (skip >= opt->step_size) {
            t += ceilf(skip / opt->step_size) * opt->step_size;
            continue;
        }
        int32_t lk[8];
        corner_links(ray, g, lk);
        const float sigma = trilerp_links(lk, g->density, 1, 0, ray->pos);
        if (opt->last_sample_opaque && t + opt->step_size > ray->tmax) ray->world_step = 1e9f;
        if (sigma > opt->sigma_thresh) {
            float col[3] = {0.f, 0.f, 0.f};
            for (int j = 0; j < D; ++j) {
                float lc = trilerp_links(lk, g->sh, D, j, ray->pos);
                col[j / B] += lc * sph[j % B];
            }
            const float pcnt = ray->world_step * sigma;
            const float weight = EXPF(log_transmit) * (1.f - EXPF(-pcnt));
            log_transmit -= pcnt;

            float tc[3], in01[3];
            for (int c = 0; c < 3; ++c) {
                const float lct = col[c] + 0.5f;
                const float clamped = maxf(lct, 0.f);
                in01[c] = (clamped == lct) ? 1.f : 0.f;
                tc[c] = clamped * grad_output[c];
            }
            float total_color = tc[0];
            total_color += tc[2];
            total_color += tc[1];

            accum -= weight * total_color;
            float curr_grad_sigma = ray->world_step * (total_color * EXPF(log_transmit) - accum);
            if (sparsity_loss > 0.f)
                curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));

            float w[8];
            for (int j = 0; j < D; ++j) {
                const int c = j / B;
                const float grad_common = weight * in01[c] * grad_output[c];
                const float curr_grad_color = sph[j % B] * grad_common;
                trilerp_weights(ray->pos, curr_grad_color, w);
                for (int q = 0; q < 8; ++q)
                    if (lk[q] >= 0) acc_sh(gr, (int64_t)lk[q] * D + j, w[q]);
            }
            trilerp_weights(ray->pos, curr_grad_sigma, w);
            for (int q = 0; q < 8; ++q)
                if (lk[q] >= 0) {
                    acc_d(gr, lk[q], w[q]);
                    if (gr->mask) gr->mask[lk[q]] = 1;
                }
            if (EXPF(log_transmit) < opt->stop_thresh) break;
        }
        t += opt->step_size;
    }
    if (accum_out) *accum_out = accum - beta_loss; /* "cancel beta loss out", cuvol.cu:375-377 */
    if (logT_out) *logT_out = log_transmit;
}

/* ---------------------------------------------------------------------------------------
 * MSI background: ray_find_bounds_bg render_util.cuh:630-676; forward cuvol.cu:386-476;
 * backward cuvol.cu:478-607; trilerp_bg_one / trilerp_backward_bg_one render_util.cuh:157-232
 * ------------------------------------------------------------------------------------- */
static void ray_find_bounds_bg(ray_t* ray, const og_grid_t* g) {
    for (int i = 0; i < 3; ++i) ray->origin[i] = fmaf(ray->origin[i], g->scaling[i], g->offset[i]);
    for (int i = 0; i < 3; ++i) ray->dir[i] *= g->scaling[i];
    const float delta_scale = rnorm3(ray->dir);
    for (int i = 0; i < 3; ++i) ray->dir[i] *= delta_scale;
    ray->world_step = delta_scale;
    for (int i = 0; i < 3; ++i) {
        const float ss = 2.f / (float)g->size[i];
        ray->origin[i] = fmaf(ray->origin[i] + 0.5f, ss, -1.f);
        ray->dir[i] = ray->dir[i] * ss;
    }
    const float inorm = rnorm3(ray->dir);
    ray->world_step *= inorm;
    for (int i = 0; i < 3; ++i) ray->dir[i] *= inorm;
}

/* _unitvec2equirect render_util.cuh:482-490 and the layer coordinate cuvol.cu:418-431 */
static void bg_locate(ray_t* ray, const og_grid_t* g, float t, float* invr_mid_out) {
    const int R = g->bg_reso, L = g->bg_nlayers;
    for (int j = 0; j < 3; ++j) ray->pos[j] = fmaf(t, ray->dir[j], ray->origin[j]);
    const float invr_mid = rnorm3(ray->pos);
    for (int j = 0; j < 3; ++j) ray->pos[j] *= invr_mid;
    const float lat = asinf(ray->pos[1]);
    const float lon = atan2f(ray->pos[0], ray->pos[2]);
    ray->pos[0] = (float)((double)(R * 2) * (0.5 + (double)lon * 0.5 * M_1_PI));
    ray->pos[1] = (float)((double)R * (0.5 - (double)lat * M_1_PI));
    ray->pos[2] = fminf(fmaxf((1.f - invr_mid) * (float)L - 0.5f, 0.f), (float)(L - 1));
    for (int j = 0; j < 3; ++j) ray->l[j] = (int)ray->pos[j];
    ray->l[0] = mini(ray->l[0], R * 2 - 1);
    ray->l[1] = mini(ray->l[1], R - 1);
    ray->l[2] = mini(ray->l[2], L - 2);
    for (int j = 0; j < 3; ++j) ray->pos[j] -= (float)ray->l[j];
    *invr_mid_out = invr_mid;
}

static void bg_links4(const ray_t* ray, const og_grid_t* g, int32_t* lk) {
    const int R = g->bg_reso;
    const int ny = ray->l[1] < (R - 1) ? (ray->l[1] + 1) : 0;
    const int nx = ray->l[0] < (2 * R - 1) ? (ray->l[0] + 1) : 0;
    lk[0] = g->bg_links[R * ray->l[0] + ray->l[1]];
    lk[1] = g->bg_links[R * ray->l[0] + ny];
    lk[2] = g->bg_links[R * nx + ray->l[1]];
    lk[3] = g->bg_links[R * nx + ny];
}

static float trilerp_bg(const ray_t* ray, const og_grid_t* g, const int32_t* lk, int idx) {
    const int L = g->bg_nlayers;
    float v[4];
    for (int q = 0; q < 4; ++q) {
        if (lk[q] >= 0) {
            const float* d = g->bg_data + ((int64_t)lk[q] * L + ray->l[2]) * 4 + idx;
            v[q] = lerpf(d[0], d[4], ray->pos[2]);
        } else v[q] = 0.f;
    }
    const float ix0 = lerpf(v[0], v[1], ray->pos[1]);
    const float ix1 = lerpf(v[2], v[3], ray->pos[1]);
    return lerpf(ix0, ix1, ray->pos[0]);
}

static void trilerp_bg_backward(const ray_t* ray, const og_grid_t* g, const int32_t* lk,
                                grads_t* gr, uint8_t* mask, float grad_out, int idx) {
    const int L = g->bg_nlayers;
    const float ay = 1.f - ray->pos[1], az = 1.f - ray->pos[2];
    float vals[4];
    float xo = (1.0f - ray->pos[0]) * grad_out;
    vals[0] = ay * xo; vals[1] = ray->pos[1] * xo;
    xo = ray->pos[0] * grad_out;
    vals[2] = ay * xo; vals[3] = ray->pos[1] * xo;
    for (int q = 0; q < 4; ++q) {
        if (lk[q] < 0) continue;
        const int64_t row = (int64_t)lk[q] * L + ray->l[2];
        acc_bg(gr, row * 4 + idx, vals[q] * az);
        acc_bg(gr, (row + 1) * 4 + idx, vals[q] * ray->pos[2]);
        if (mask) { mask[row] = 1; mask[row + 1] = 1; }
    }
}

static void render_background_forward(const og_grid_t* g, const og_opt_t* opt, ray_t* ray,
                                      float log_transmit, float* out, og_counts_t* cnt) {
    csi_t c = csi_make(ray->origin, ray->dir);
    float cr[3] = {ray->origin[1] * ray->dir[2] - ray->origin[2] * ray->dir[1],
                   ray->origin[2] * ray->dir[0] - ray->origin[0] * ray->dir[2],
                   ray->origin[0] * ray->dir[1] - ray->origin[1] * ray->dir[0]};
    const float inner_radius = fmaxf(norm3(cr) + 1e-3f, 1.f);
    float t, invr_last = 1.f / inner_radius;
    const int n_steps = (int)((float)g->bg_nlayers / opt->step_size) + 2;
    float outv[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < n_steps; ++i) {
        const float r = (float)((double)n_steps / ((double)(n_steps - i) - 0.5));
        if (r < inner_radius || !csi_intersect(&c, r, &t, 0)) continue;
        float invr_mid;
        bg_locate(ray, g, t, &invr_mid);
        int32_t lk[4];
        bg_links4(ray, g, lk);
        const float sigma = trilerp_bg(ray, g, lk, 3);
        if (sigma > 0.f) {
            if (cnt) cnt->bg_act++;
            const float pcnt = (invr_last - invr_mid) * ray->world_step * sigma;
            const float weight = EXPF(log_transmit) * (1.f - EXPF(-pcnt));
            log_transmit -= pcnt;
            for (int k = 0; k < 3; ++k) {
                const float color = trilerp_bg(ray, g, lk, k) * SH_C0;
                outv[k] += weight * fmaxf(color + 0.5f, 0.f);
            }
            if (EXPF(log_transmit) < opt->stop_thresh) break;
        }
        invr_last = invr_mid;
    }
    for (int k = 0; k < 3; ++k) out[k] += outv[k] + EXPF(log_transmit) * opt->background_brightness;
}

static void render_background_backward(const og_grid_t* g, const og_opt_t* opt, ray_t* ray,
                                       const float* grad_output, float log_transmit, float accum,
                                       float sparsity_loss, grads_t* gr) {
    csi_t c = csi_make(ray->origin, ray->dir);
    const int n_steps = (int)((float)g->bg_nlayers / opt->step_size) + 2;
    float cr[3] = {ray->origin[1] * ray->dir[2] - ray->origin[2] * ray->dir[1],
                   ray->origin[2] * ray->dir[0] - ray->origin[0] * ray->dir[2],
                   ray->origin[0] * ray->dir[1] - ray->origin[1] * ray->dir[0]};
    const float inner_radius = fmaxf(norm3(cr) + 1e-3f, 1.f);
    float t, invr_last = 1.f / inner_radius;
    for (int i = 0; i < n_steps; ++i) {
        const float r = (float)((double)n_steps / ((double)(n_steps - i) - 0.5));
        if (r < inner_radius || !csi_intersect(&c, r, &t, 0)) continue;
        float invr_mid;
        bg_locate(ray, g, t, &invr_mid);
        int32_t lk[4];
        bg_links4(ray, g, lk);
        const float sigma = trilerp_bg(ray, g, lk, 3);
        if (sigma > 0.f) {
            float total_color = 0.f;
            const float pcnt = ray->world_step * (invr_last - invr_mid) * sigma;
            const float weight = EXPF(log_transmit) * (1.f - EXPF(-pcnt));
            log_transmit -= pcnt;
            for (int k = 0; k < 3; ++k) {
                const float color = trilerp_bg(ray, g, lk, k) * SH_C0 + 0.5f;
                total_color += fmaxf(color, 0.f) * grad_output[k];
                if (color > 0.f) {
                    const float curr_grad_color = SH_C0 * weight * grad_output[k];
                    trilerp_bg_backward(ray, g, lk, gr, NULL, curr_grad_color, k);
                }
            }
            accum -= weight * total_color;
            float curr_grad_sigma = ray->world_step * (invr_last - invr_mid) *
                                    (total_color * EXPF(log_transmit) - accum);
            if (sparsity_loss > 0.f)
                curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
            trilerp_bg_backward(ray, g, lk, gr, gr->mask_bg, curr_grad_sigma, 3);
            if (EXPF(log_transmit) < opt->stop_thresh) break;
        }
        invr_last = invr_mid;
    }
}

/* ---------------------------------------------------------------------------------------
 * camera rays: cam2world_ray / world2ndc, render_util.cuh:504-544
 * ------------------------------------------------------------------------------------- */
static void cam2world_ray(int ix, int iy, const og_cam_t* cam, float* dir, float* origin) {
    float x = ((float)ix + 0.5f - cam->cx) / cam->fx;
    float y = ((float)iy + 0.5f - cam->cy) / cam->fy;
    float z = sqrtf((float)((double)(x * x + y * y) + 1.0));
    x /= z; y /= z; z = 1.0f / z;
    const float* m = cam->c2w;
    dir[0] = m[0] * x + m[1] * y + m[2] * z;
    dir[1] = m[4] * x + m[5] * y + m[6] * z;
    dir[2] = m[8] * x + m[9] * y + m[10] * z;
    origin[0] = m[3]; origin[1] = m[7]; origin[2] = m[11];
    if (cam->ndc_coeffx > 0.f) {
        const float near = 1.f;
        const float t = (near - origin[2]) / dir[2];
        for (int i = 0; i < 3; ++i) origin[i] = fmaf(t, dir[i], origin[i]);
        dir[0] = cam->ndc_coeffx * (dir[0] / dir[2] - origin[0] / origin[2]);
        dir[1] = cam->ndc_coeffy * (dir[1] / dir[2] - origin[1] / origin[2]);
        dir[2] = 2 * near / origin[2];
        origin[0] = cam->ndc_coeffx * (origin[0] / origin[2]);
        origin[1] = cam->ndc_coeffy * (origin[1] / origin[2]);
        origin[2] = 1 - 2 * near / origin[2];
        const float rn = rnorm3(dir);
        dir[0] *= rn; dir[1] *= rn; dir[2] *= rn;
    }
}

void og_gen_rays(const og_cam_t* cam, float* origins, float* dirs) {
    for (int iy = 0; iy < cam->height; ++iy)
        for (int ix = 0; ix < cam->width; ++ix) {
            const int64_t r = (int64_t)iy * cam->width + ix;
            cam2world_ray(ix, iy, cam, dirs + 3 * r, origins + 3 * r);
        }
}

/* ---------------------------------------------------------------------------------------
 * public render entry points
 * ------------------------------------------------------------------------------------- */
static void set_ray(ray_t* r, const float* o, const float* d) {
    for (int i = 0; i < 3; ++i) { r->origin[i] = o[i]; r->dir[i] = d[i]; }
}

static void render_one(const og_grid_t* g, const og_opt_t* o, const float* origin,
                       const float* dir, float* rgb, float* logT_out, og_counts_t* counts) {
    ray_t ray;
    float sph[9];
    set_ray(&ray, origin, dir);
    calc_sh(g->basis_dim, dir, sph); /* raw input direction, cuvol.cu:632-635 */
    ray_find_bounds(&ray, g, o);
    float logT = 0.f;
    trace_ray_fwd(g, o, &ray, sph, rgb, &logT, counts);
    if (g->bg_nlayers > 0 && !(logT < -25.f)) {
        ray_t bray;
        set_ray(&bray, origin, dir);
        ray_find_bounds_bg(&bray, g);
        render_background_forward(g, o, &bray, logT, rgb, counts);
    }
    if (logT_out) *logT_out = logT;
}

typedef struct {
    const og_grid_t* g; const og_opt_t* o; const float* origins; const float* dirs;
    float* rgb_out; float* log_transmit; og_counts_t counts;
} fwd_ctx_t;
static void fwd_range(int64_t b, int64_t e, void* vctx) {
    fwd_ctx_t* c = (fwd_ctx_t*)vctx;
    og_counts_t loc;
    memset(&loc, 0, sizeof(loc));
    for (int64_t r = b; r < e; ++r)
        render_one(c->g, c->o, c->origins + 3 * r, c->dirs + 3 * r, c->rgb_out + 3 * r,
                   c->log_transmit ? c->log_transmit + r : NULL, &loc);
    __atomic_fetch_add(&c->counts.s_skip, loc.s_skip, __ATOMIC_RELAXED);
    __atomic_fetch_add(&c->counts.s_vis, loc.s_vis, __ATOMIC_RELAXED);
    __atomic_fetch_add(&c->counts.s_act, loc.s_act, __ATOMIC_RELAXED);
    __atomic_fetch_add(&c->counts.bg_act, loc.bg_act, __ATOMIC_RELAXED);
}

void og_volume_render(const og_grid_t* g, const og_opt_t* o, const float* origins,
                      const float* dirs, int64_t Q, float* rgb_out, float* log_transmit,
                      og_counts_t* counts) {
    fwd_ctx_t c = {g, o, origins, dirs, rgb_out, log_transmit, {0, 0, 0, 0}};
    parallel_for(Q, fwd_range, &c);
    if (counts) *counts = c.counts;
}

void og_volume_render_image(const og_grid_t* g, const og_cam_t* cam, const og_opt_t* o,
                            float* out) {
    for (int iy = 0; iy < cam->height; ++iy)
        for (int ix = 0; ix < cam->width; ++ix) {
            float origin[3], dir[3];
            cam2world_ray(ix, iy, cam, dir, origin);
            render_one(g, o, origin, dir, out + 3 * ((int64_t)iy * cam->width + ix), NULL, NULL);
        }
}

static void grads_begin_f(const og_grid_t* g, grads_t* gr, float* grad_density, float* grad_sh,
                          float* grad_bg, uint8_t* mask, uint8_t* mask_bg);
static void grads_begin(const og_grid_t* g, grads_t* gr, int want_bg, uint8_t* mask,
                        uint8_t* mask_bg) {
    const int D = 3 * g->basis_dim;
    gr->fgd = gr->fgsh = gr->fgbg = NULL;
    gr->gd = (double*)calloc((size_t)g->n_rows, sizeof(double));
    gr->gsh = (double*)calloc((size_t)g->n_rows * D, sizeof(double));
    gr->gbg = (want_bg && g->bg_nlayers > 0)
                  ? (double*)calloc((size_t)g->bg_rows * g->bg_nlayers * 4, sizeof(double))
                  : NULL;
    gr->mask = mask;
    gr->mask_bg = mask_bg;
}
/* bench mode: accumulate in float directly into the caller's arrays (what the reference's
 * atomicAdd does), so the timed call carries no O(N) staging pass. */
static void grads_begin_f(const og_grid_t* g, grads_t* gr, float* grad_density, float* grad_sh,
                          float* grad_bg, uint8_t* mask, uint8_t* mask_bg) {
    if (!g_accum_float) { grads_begin(g, gr, grad_bg != NULL, mask, mask_bg); return; }
    gr->gd = gr->gsh = gr->gbg = NULL;
    gr->fgd = grad_density; gr->fgsh = grad_sh; gr->fgbg = (g->bg_nlayers > 0) ? grad_bg : NULL;
    gr->mask = mask; gr->mask_bg = mask_bg;
}
static void grads_end(const og_grid_t* g, grads_t* gr, float* grad_density, float* grad_sh,
                      float* grad_bg) {
    const int D = 3 * g->basis_dim;
    if (gr->fgd) return;
    for (int64_t i = 0; i < g->n_rows; ++i) grad_density[i] = (float)((double)grad_density[i] + gr->gd[i]);
    for (int64_t i = 0; i < g->n_rows * D; ++i) grad_sh[i] = (float)((double)grad_sh[i] + gr->gsh[i]);
    if (gr->gbg && grad_bg) {
        const int64_t n = g->bg_rows * g->bg_nlayers * 4;
        for (int64_t i = 0; i < n; ++i) grad_bg[i] = (float)((double)grad_bg[i] + gr->gbg[i]);
    }
    free(gr->gd); free(gr->gsh); free(gr->gbg);
}

/* One ray of render_ray_backward_kernel (+ render_background_backward_kernel),
 * cuvol.cu:692-773, 818-860.  fused: grad_in is rgb_gt and the seed is (out-gt)*2/(3*norm). */
static void backward_one(const og_grid_t* g, const og_opt_t* o, const float* origin,
                         const float* dir, const float* grad_in, const float* color_cache,
                         int fused, int64_t norm_rays, const float* logT_fwd, float beta_loss,
                         float sparsity_loss, grads_t* gr) {
    ray_t ray;
    float sph[9];
    set_ray(&ray, origin, dir);
    calc_sh(g->basis_dim, dir, sph);
    ray_find_bounds(&ray, g, o);
    float gout[3];
    if (fused) {
        const float norm_factor = 2.f / (float)(3 * (int)norm_rays);
        for (int i = 0; i < 3; ++i) gout[i] = (color_cache[i] - grad_in[i]) * norm_factor;
    } else {
        for (int i = 0; i < 3; ++i) gout[i] = grad_in[i];
    }
    const int has_bg = g->bg_nlayers > 0;
    float accum = 0.f, logT_bwd = 0.f;
    trace_ray_bwd(g, o, &ray, sph, gout, color_cache,
                  (fused && beta_loss > 0.f && logT_fwd) ? *logT_fwd : 0.f,
                  fused ? beta_loss : 0.f, fused ? sparsity_loss : 0.f, gr, &accum, &logT_bwd);
    if (has_bg && (gr->gbg || gr->fgbg)) {
        /* fused: bg kernels read the FORWARD log_transmit (cuvol.cu:1143); non-fused: the
         * value re-derived by the backward march (cuvol.cu:1048,1059). */
        const float logT = fused ? *logT_fwd : logT_bwd;
        if (!(logT < -25.f)) {
            ray_t bray;
            set_ray(&bray, origin, dir);
            ray_find_bounds_bg(&bray, g);
            render_background_backward(g, o, &bray, gout, logT, accum,
                                       fused ? sparsity_loss : 0.f, gr);
        }
    }
}

typedef struct {
    const og_grid_t* g; const og_opt_t* o; const float* origins; const float* dirs;
    const float* grad_in; const float* color_cache; int fused; int64_t norm_rays;
    const float* logT; float beta; float sparsity; grads_t* gr;
} bwd_ctx_t;
static void bwd_range(int64_t b, int64_t e, void* vctx) {
    bwd_ctx_t* c = (bwd_ctx_t*)vctx;
    for (int64_t r = b; r < e; ++r)
        backward_one(c->g, c->o, c->origins + 3 * r, c->dirs + 3 * r, c->grad_in + 3 * r,
                     c->color_cache + 3 * r, c->fused, c->norm_rays, c->logT ? c->logT + r : NULL,
                     c->beta, c->sparsity, c->gr);
}

void og_volume_render_backward(const og_grid_t* g, const og_opt_t* o, const float* origins,
                               const float* dirs, int64_t Q, const float* grad_out,
                               const float* color_cache, float* grad_density, float* grad_sh,
                               float* grad_bg, uint8_t* mask, uint8_t* mask_bg) {
    grads_t gr;
    grads_begin_f(g, &gr, grad_density, grad_sh, grad_bg, mask, mask_bg);
    bwd_ctx_t c = {g, o, origins, dirs, grad_out, color_cache, 0, Q, NULL, 0.f, 0.f, &gr};
    parallel_for(Q, bwd_range, &c);
    grads_end(g, &gr, grad_density, grad_sh, grad_bg);
}

void og_volume_render_fused(const og_grid_t* g, const og_opt_t* o, const float* origins,
                            const float* dirs, int64_t Q, int64_t norm_rays,
                            const float* rgb_gt, float beta_loss, float sparsity_loss,
                            float* rgb_out, float* grad_density, float* grad_sh, float* grad_bg,
                            uint8_t* mask, uint8_t* mask_bg, og_counts_t* counts) {
    float* logT = (float*)malloc(sizeof(float) * (size_t)(Q > 0 ? Q : 1));
    og_volume_render(g, o, origins, dirs, Q, rgb_out, logT, counts);
    grads_t gr;
    grads_begin_f(g, &gr, grad_density, grad_sh, grad_bg, mask, mask_bg);
    const float beta = beta_loss / (float)norm_rays; /* cuvol.cu:1127 */
    bwd_ctx_t c = {g, o, origins, dirs, rgb_gt, rgb_out, 1, norm_rays, logT, beta, sparsity_loss, &gr};
    parallel_for(Q, bwd_range, &c);
    grads_end(g, &gr, grad_density, grad_sh, grad_bg);
    free(logT);
}

/* ---------------------------------------------------------------------------------------
 * nvol backend (Neural-Volumes compositing): trace_ray_nvol, render_lerp_kernel_nvol.cu:30-111;
 * trace_ray_nvol_backward :113-296; launchers :406-502.  No MSI background, no beta loss.
 * ------------------------------------------------------------------------------------- */
static void trace_ray_nvol_fwd(const og_grid_t* g, const og_opt_t* opt, ray_t* ray, const float* sph,
                               float* out) {
    const int B = g->basis_dim, D = 3 * B;
    const float bgb = (g->bg_nlayers == 0) ? opt->background_brightness : 0.f;
    if (ray->tmin > ray->tmax) {
        for (int c = 0; c < 3; ++c) out[c] = bgb;
        return;
    }
    float t = ray->tmin, total_alpha = 0.f;
    float outv[3] = {0.f, 0.f, 0.f};
    while (t <= ray->tmax) {
        ray_locate(ray, g, t);
        const float skip = compute_skip_dist(ray, g);
        if (skip >= opt->step_size) {
            t += ceilf(skip / opt->step_size) * opt->step_size;
            continue;
        }
        int32_t lk[8];
        corner_links(ray, g, lk);
        const float sigma = trilerp_links(lk, g->density, 1, 0, ray->pos);
        if (sigma > opt->sigma_thresh) {
            float col[3] = {0.f, 0.f, 0.f};
            for (int j = 0; j < D; ++j)
                col[j / B] += trilerp_links(lk, g->sh, D, j, ray->pos) * sph[j % B];
            const float nta = minf(total_alpha + 1.f - EXPF(-ray->world_step * sigma), 1.f);
            const float weight = nta - total_alpha;
            total_alpha = nta;
            for (int c = 0; c < 3; ++c) outv[c] += weight * maxf(col[c] + 0.5f, 0.f);
            if (total_alpha >= 1.f) break;
        }
        t += opt->step_size;
    }
    for (int c = 0; c < 3; ++c) out[c] = outv[c] + (1.f - total_alpha) * bgb;
}

static void trace_ray_nvol_bwd(const og_grid_t* g, const og_opt_t* opt, ray_t* ray, const float* sph,
                               const float* gout, float sparsity_loss, grads_t* gr) {
    const int B = g->basis_dim, D = 3 * B;
    if (ray->tmin > ray->tmax) return;
    float t = ray->tmin, total_alpha = 0.f, last_total_color = 0.f;
    float w[8];
    while (t <= ray->tmax) { /* pass 1: active samples (:137-241) */
        ray_locate(ray, g, t);
        const float skip = compute_skip_dist(ray, g);
        if (skip >= opt->step_size) {
            t += ceilf(skip / opt->step_size) * opt->step_size;
            continue;
        }
        int32_t lk[8];
        corner_links(ray, g, lk);
        const float sigma = trilerp_links(lk, g->density, 1, 0, ray->pos);
        if (sigma > opt->sigma_thresh) {
            float col[3] = {0.f, 0.f, 0.f};
            for (int j = 0; j < D; ++j)
                col[j / B] += trilerp_links(lk, g->sh, D, j, ray->pos) * sph[j % B];
            const float curr_transmit = EXPF(-ray->world_step * sigma);
            const float nta = minf(total_alpha + 1.f - curr_transmit, 1.f);
            const float weight = nta - total_alpha;
            const int not_last = nta < 1.f;
            total_alpha = nta;
            float tc[3], in01[3];
            for (int c = 0; c < 3; ++c) {
                const float lct = col[c] + 0.5f;
                const float clamped = maxf(lct, 0.f);
                in01[c] = (clamped == lct) ? 1.f : 0.f;
                tc[c] = clamped * gout[c];
            }
            float total_color = tc[0];
            total_color += tc[2];
            total_color += tc[1];
            for (int j = 0; j < D; ++j) {
                const int c = j / B;
                const float grad_common = weight * in01[c] * gout[c];
                trilerp_weights(ray->pos, sph[j % B] * grad_common, w);
                for (int q = 0; q < 8; ++q)
                    if (lk[q] >= 0) acc_sh(gr, (int64_t)lk[q] * D + j, w[q]);
            }
            if (not_last) {
                float curr_grad_sigma = ray->world_step * curr_transmit * total_color;
                if (sparsity_loss > 0.f)
                    curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
                trilerp_weights(ray->pos, curr_grad_sigma, w);
                for (int q = 0; q < 8; ++q)
                    if (lk[q] >= 0) {
                        acc_d(gr, lk[q], w[q]);
                        if (gr->mask) gr->mask[lk[q]] = 1;
                    }
            } else {
                ray->tmax = t;
                last_total_color = total_color;
                break;
            }
        }
        t += opt->step_size;
    }
    if (total_alpha < 1.f) /* never saturated (:243-247) */
        last_total_color = opt->background_brightness * (gout[0] + gout[1] + gout[2]);
    if (last_total_color == 0.f) return;
    t = ray->tmin;
    total_alpha = 0.f;
    while (t <= ray->tmax) { /* pass 2: every visited sample before saturation (:248-293) */
        ray_locate(ray, g, t);
        const float skip = compute_skip_dist(ray, g);
        if (skip >= opt->step_size) {
            t += ceilf(skip / opt->step_size) * opt->step_size;
            continue;
        }
        int32_t lk[8];
        corner_links(ray, g, lk);
        const float sigma = trilerp_links(lk, g->density, 1, 0, ray->pos);
        const float curr_transmit = EXPF(-ray->world_step * sigma);
        total_alpha = minf(total_alpha + 1.f - curr_transmit, 1.f);
        if (total_alpha >= 1.f) break;
        trilerp_weights(ray->pos, -ray->world_step * curr_transmit * last_total_color, w);
        for (int q = 0; q < 8; ++q)
            if (lk[q] >= 0) {
                acc_d(gr, lk[q], w[q]);
                if (gr->mask) gr->mask[lk[q]] = 1;
            }
        t += opt->step_size;
    }
}

typedef struct {
    const og_grid_t* g; const og_opt_t* o; const float* origins; const float* dirs;
    float* rgb_out; const float* grad_in; const float* color_cache; int fused; int64_t norm_rays;
    float sparsity; grads_t* gr;
} nvol_ctx_t;
static void nvol_fwd_range(int64_t b, int64_t e, void* vctx) {
    nvol_ctx_t* c = (nvol_ctx_t*)vctx;
    for (int64_t r = b; r < e; ++r) {
        ray_t ray;
        float sph[9];
        set_ray(&ray, c->origins + 3 * r, c->dirs + 3 * r);
        calc_sh(c->g->basis_dim, c->dirs + 3 * r, sph);
        ray_find_bounds(&ray, c->g, c->o);
        trace_ray_nvol_fwd(c->g, c->o, &ray, sph, c->rgb_out + 3 * r);
    }
}
static void nvol_bwd_range(int64_t b, int64_t e, void* vctx) {
    nvol_ctx_t* c = (nvol_ctx_t*)vctx;
    for (int64_t r = b; r < e; ++r) {
        ray_t ray;
        float sph[9], gout[3];
        set_ray(&ray, c->origins + 3 * r, c->dirs + 3 * r);
        calc_sh(c->g->basis_dim, c->dirs + 3 * r, sph);
        ray_find_bounds(&ray, c->g, c->o);
        if (c->fused) { /* :364-370 */
            const float norm_factor = 2.f / (float)(3 * (int)c->norm_rays);
            for (int i = 0; i < 3; ++i)
                gout[i] = (c->color_cache[3 * r + i] - c->grad_in[3 * r + i]) * norm_factor;
        } else {
            for (int i = 0; i < 3; ++i) gout[i] = c->grad_in[3 * r + i];
        }
        trace_ray_nvol_bwd(c->g, c->o, &ray, sph, gout, c->fused ? c->sparsity : 0.f, c->gr);
    }
}

void og_volume_render_nvol(const og_grid_t* g, const og_opt_t* o, const float* origins,
                           const float* dirs, int64_t Q, float* rgb_out) {
    nvol_ctx_t c = {g, o, origins, dirs, rgb_out, NULL, NULL, 0, Q, 0.f, NULL};
    parallel_for(Q, nvol_fwd_range, &c);
}

void og_volume_render_nvol_backward(const og_grid_t* g, const og_opt_t* o, const float* origins,
                                    const float* dirs, int64_t Q, const float* grad_out,
                                    const float* color_cache, float* grad_density, float* grad_sh,
                                    uint8_t* mask) {
    grads_t gr;
    grads_begin_f(g, &gr, grad_density, grad_sh, NULL, mask, NULL);
    nvol_ctx_t c = {g, o, origins, dirs, NULL, grad_out, color_cache, 0, Q, 0.f, &gr};
    parallel_for(Q, nvol_bwd_range, &c);
    grads_end(g, &gr, grad_density, grad_sh, NULL);
}

void og_volume_render_nvol_fused(const og_grid_t* g, const og_opt_t* o, const float* origins,
                                 const float* dirs, int64_t Q, int64_t norm_rays,
                                 const float* rgb_gt, float sparsity_loss, float* rgb_out,
                                 float* grad_density, float* grad_sh, uint8_t* mask) {
    og_volume_render_nvol(g, o, origins, dirs, Q, rgb_out);
    grads_t gr;
    grads_begin_f(g, &gr, grad_density, grad_sh, NULL, mask, NULL);
    nvol_ctx_t c = {g, o, origins, dirs, NULL, rgb_gt, rgb_out, 1, norm_rays, sparsity_loss, &gr};
    parallel_for(Q, nvol_bwd_range, &c);
    grads_end(g, &gr, grad_density, grad_sh, NULL);
}

/* ---------------------------------------------------------------------------------------
 * depth: trace_ray_expected_term cuvol.cu:122-177, trace_ray_sigma_thresh cuvol.cu:180-225
 * ------------------------------------------------------------------------------------- */
void og_render_expected_term(const og_grid_t* g, const og_opt_t* opt, const float* origins,
                             const float* dirs, int64_t Q, float* out) {
    for (int64_t r = 0; r < Q; ++r) {
        ray_t ray;
        set_ray(&ray, origins + 3 * r, dirs + 3 * r);
        ray_find_bounds(&ray, g, opt);
        if (ray.tmin > ray.tmax) { out[r] = 0.f; continue; }
        float t = ray.tmin, outv = 0.f, log_transmit = 0.f;
        while (t <= ray.tmax) {
            ray_locate(&ray, g, t);
            const float skip = compute_skip_dist(&ray, g);
            if (skip >= opt->step_size) {
                t += ceilf(skip / opt->step_size) * opt->step_size;
                continue;
            }
            int32_t lk[8];
            corner_links(&ray, g, lk);
            const float sigma = trilerp_links(lk, g->density, 1, 0, ray.pos);
            if (sigma > opt->sigma_thresh) {
                const float pcnt = ray.world_step * sigma;
                const float weight = EXPF(log_transmit) * (1.f - EXPF(-pcnt));
                log_transmit -= pcnt;
                outv += weight * (t / opt->step_size) * ray.world_step;
                if (EXPF(log_transmit) < opt->stop_thresh) break;
            }
            t += opt->step_size;
        }
        out[r] = outv;
    }
}

void og_render_sigma_thresh(const og_grid_t* g, const og_opt_t* opt, const float* origins,
                            const float* dirs, int64_t Q, float sigma_thresh, float* out) {
    for (int64_t r = 0; r < Q; ++r) {
        ray_t ray;
        set_ray(&ray, origins + 3 * r, dirs + 3 * r);
        ray_find_bounds(&ray, g, opt);
        out[r] = 0.f;
        if (ray.tmin > ray.tmax) continue;
        float t = ray.tmin;
        while (t <= ray.tmax) {
            ray_locate(&ray, g, t);
            const float skip = compute_skip_dist(&ray, g);
            if (skip >= opt->step_size) {
                t += ceilf(skip / opt->step_size) * opt->step_size;
                continue;
            }
            int32_t lk[8];
            corner_links(&ray, g, lk);
            const float sigma = trilerp_links(lk, g->density, 1, 0, ray.pos);
            if (sigma > sigma_thresh) {
                out[r] = (t / opt->step_size) * ray.world_step;
                break;
            }
            t += opt->step_size;
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * accel_dist_prop: misc_kernel.cu:109-181 (kernels), 356-392 (host)
 * ------------------------------------------------------------------------------------- */
#define DIV2_CEIL(x) ((((x) - 1) >> 1) + 1)

void og_accel_dist_prop(int32_t* links, int32_t X, int32_t Y, int32_t Z) {
    int64_t req = 0;
    {
        int64_t sx = X, sy = Y, sz = Z;
        while (sx > 1 && sy > 1 && sz > 1) {
            sx = DIV2_CEIL(sx); sy = DIV2_CEIL(sy); sz = DIV2_CEIL(sz);
            req += sx * sy * sz;
        }
    }
    uint8_t* tmp = (uint8_t*)calloc((size_t)(req > 0 ? req : 1), 1);
    for (int x0 = 0; x0 < X; ++x0)
        for (int y0 = 0; y0 < Y; ++y0)
            for (int z0 = 0; z0 < Z; ++z0) {
                if (links[((int64_t)x0 * Y + y0) * Z + z0] < 0) continue;
                int x = x0, y = y0, z = z0, sx = X, sy = Y, sz = Z;
                uint8_t* base = tmp;
                while (sx > 1 && sy > 1 && sz > 1) {
                    x >>= 1; y >>= 1; z >>= 1;
                    sx = DIV2_CEIL(sx); sy = DIV2_CEIL(sy); sz = DIV2_CEIL(sz);
                    base[(int64_t)x * sy * sz + (int64_t)y * sz + z] = 1;
                    base += (int64_t)sx * sy * sz;
                }
            }
    for (int x0 = 0; x0 < X; ++x0)
        for (int y0 = 0; y0 < Y; ++y0)
            for (int z0 = 0; z0 < Z; ++z0) {
                int32_t* val = &links[((int64_t)x0 * Y + y0) * Z + z0];
                if (*val >= 0) continue;
                int x = x0, y = y0, z = z0, sx = X, sy = Y, sz = Z;
                const uint8_t* base = tmp;
                int result = -1;
                while (sx > 1 && sy > 1 && sz > 1) {
                    x >>= 1; y >>= 1; z >>= 1;
                    sx = DIV2_CEIL(sx); sy = DIV2_CEIL(sy); sz = DIV2_CEIL(sz);
                    if (base[(int64_t)x * sy * sz + (int64_t)y * sz + z]) break;
                    result -= 1;
                    base += (int64_t)sx * sy * sz;
                }
                *val = result;
            }
    free(tmp);
}

/* dilate: misc_kernel.cu:20-48 */
void og_dilate(const uint8_t* grid, int32_t X, int32_t Y, int32_t Z, uint8_t* out) {
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z) {
                uint8_t v = 0;
                for (int dx = -1; dx <= 1; ++dx)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dz = -1; dz <= 1; ++dz) {
                            const int xx = mini(maxi(x + dx, 0), X - 1);
                            const int yy = mini(maxi(y + dy, 0), Y - 1);
                            const int zz = mini(maxi(z + dz, 0), Z - 1);
                            v |= grid[((int64_t)xx * Y + yy) * Z + zz];
                        }
                out[((int64_t)x * Y + y) * Z + z] = v ? 1 : 0;
            }
}

/* grid_weight_render: grid_trace_ray misc_kernel.cu:184-280 */
void og_grid_weight_render(const float* data, int32_t X, int32_t Y, int32_t Z,
                           const og_cam_t* cam, float step_size, float stop_thresh,
                           int32_t last_sample_opaque, const float* offset,
                           const float* scaling, float* grid_weight) {
    const int size[3] = {X, Y, Z};
    const int64_t stride0 = (int64_t)Y * Z, stride1 = Z;
    for (int iy = 0; iy < cam->height; ++iy)
        for (int ix = 0; ix < cam->width; ++ix) {
            float o[3], d[3];
            cam2world_ray(ix, iy, cam, d, o);
            for (int i = 0; i < 3; ++i) o[i] = fmaf(o[i], scaling[i], offset[i]);
            for (int i = 0; i < 3; ++i) d[i] *= scaling[i];
            const float delta_scale = rnorm3(d);
            for (int i = 0; i < 3; ++i) d[i] *= delta_scale;
            const float world_step = delta_scale * step_size;
            float t = 0.f, tmax = 2e3f;
            for (int i = 0; i < 3; ++i) {
                const float invdir = 1.0f / d[i];
                const float t1 = (-0.5f - o[i]) * invdir;
                const float t2 = ((float)size[i] - 0.5f - o[i]) * invdir;
                if (d[i] != 0.f) {
                    t = maxf(t, minf(t1, t2));
                    tmax = minf(tmax, maxf(t1, t2));
                }
            }
            if (t > tmax) continue;
            float logI = 0.f;
            while (t <= tmax) {
                float pos[3];
                int32_t l[3];
                for (int j = 0; j < 3; ++j) {
                    pos[j] = fmaf(t, d[j], o[j]); /* o + t*d, contracted by nvcc */
                    pos[j] = minf(maxf(pos[j], 0.f), (float)size[j] - 1.f);
                    l[j] = mini((int32_t)pos[j], size[j] - 2);
                    pos[j] -= (float)l[j];
                }
                const int64_t idx = l[0] * stride0 + l[1] * stride1 + l[2];
                const float* s000 = data + idx;
                const float* s100 = s000 + stride0;
                const float ix0y0 = lerpf(s000[0], s000[1], pos[2]);
                const float ix0y1 = lerpf(s000[stride1], s000[stride1 + 1], pos[2]);
                const float ix1y0 = lerpf(s100[0], s100[1], pos[2]);
                const float ix1y1 = lerpf(s100[stride1], s100[stride1 + 1], pos[2]);
                const float ix0 = lerpf(ix0y0, ix0y1, pos[1]);
                const float ix1 = lerpf(ix1y0, ix1y1, pos[1]);
                float sigma = lerpf(ix0, ix1, pos[0]);
                if (last_sample_opaque && t + step_size > tmax) {
                    sigma += 1e9f;
                    logI = 0.f;
                }
                if (sigma > 1e-8f) {
                    const float log_att = -world_step * sigma;
                    const float weight = EXPF(logI) * (1.f - EXPF(log_att));
                    logI += log_att;
                    float* w000 = grid_weight + idx;
                    float* w100 = w000 + stride0;
                    w000[0] = fmaxf(w000[0], weight);
                    w000[1] = fmaxf(w000[1], weight);
                    w000[stride1] = fmaxf(w000[stride1], weight);
                    w000[stride1 + 1] = fmaxf(w000[stride1 + 1], weight);
                    w100[0] = fmaxf(w100[0], weight);
                    w100[1] = fmaxf(w100[1], weight);
                    w100[stride1] = fmaxf(w100[stride1], weight);
                    w100[stride1 + 1] = fmaxf(w100[stride1 + 1], weight);
                    if (EXPF(logI) < stop_thresh) break;
                }
                t += step_size;
            }
        }
}

/* ---------------------------------------------------------------------------------------
 * sample_grid: svox2_kernel.cu:10-78 ; backward :81-176
 * ------------------------------------------------------------------------------------- */
static void sample_locate(const og_grid_t* g, const float* p, int32_t* l, float* pos,
                          int32_t* lk) {
    for (int i = 0; i < 3; ++i) {
        float v = fmaf(p[i], g->scaling[i], g->offset[i]);
        v = fminf(fmaxf(v, 0.f), (float)g->size[i] - 1.f);
        l[i] = mini((int32_t)v, g->size[i] - 2);
        pos[i] = v - (float)l[i];
    }
    ray_t tmp;
    for (int i = 0; i < 3; ++i) tmp.l[i] = l[i];
    corner_links(&tmp, g, lk);
}

void og_sample_grid(const og_grid_t* g, const float* points, int64_t P, float* out_density,
                    float* out_sh) {
    const int D = 3 * g->basis_dim;
    for (int64_t p = 0; p < P; ++p) {
        int32_t l[3], lk[8];
        float pos[3];
        sample_locate(g, points + 3 * p, l, pos, lk);
        out_density[p] = trilerp_links(lk, g->density, 1, 0, pos);
        if (out_sh)
            for (int j = 0; j < D; ++j) out_sh[p * D + j] = trilerp_links(lk, g->sh, D, j, pos);
    }
}

void og_sample_grid_backward(const og_grid_t* g, const float* points, int64_t P,
                             const float* grad_out_density, const float* grad_out_sh,
                             float* grad_density, float* grad_sh) {
    const int D = 3 * g->basis_dim;
    grads_t gr;
    grads_begin(g, &gr, 0, NULL, NULL);
    for (int64_t p = 0; p < P; ++p) {
        int32_t l[3], lk[8];
        float pos[3], w[8];
        sample_locate(g, points + 3 * p, l, pos, lk);
        trilerp_weights(pos, grad_out_density[p], w);
        for (int q = 0; q < 8; ++q)
            if (lk[q] >= 0) gr.gd[lk[q]] += (double)w[q];
        if (grad_out_sh && grad_sh)
            for (int j = 0; j < D; ++j) {
                trilerp_weights(pos, grad_out_sh[p * D + j], w);
                for (int q = 0; q < 8; ++q)
                    if (lk[q] >= 0) gr.gsh[(int64_t)lk[q] * D + j] += (double)w[q];
            }
    }
    if (!grad_sh) { free(gr.gsh); gr.gsh = NULL; }
    for (int64_t i = 0; i < g->n_rows; ++i) grad_density[i] = (float)((double)grad_density[i] + gr.gd[i]);
    if (grad_sh)
        for (int64_t i = 0; i < g->n_rows * D; ++i) grad_sh[i] = (float)((double)grad_sh[i] + gr.gsh[i]);
    free(gr.gd); free(gr.gsh); free(gr.gbg);
}

/* ---------------------------------------------------------------------------------------
 * TV: loss_kernel.cu (calculate_ray_scale :22-52; tv_kernel :64-109; tv_grad_kernel :111-176;
 * tv_grad_sparse_kernel :178-246; msi_tv_grad_sparse_kernel :248-333)
 * ------------------------------------------------------------------------------------- */
float og_tv(const int32_t* links, int32_t X, int32_t Y, int32_t Z, const float* data,
            int32_t ddim, int32_t start_dim, int32_t end_dim, int32_t ignore_edge) {
    const float sc[3] = {(float)X * (1.f / 256.f), (float)Y * (1.f / 256.f), (float)Z * (1.f / 256.f)};
    const int64_t nl = (int64_t)(X - 1) * (Y - 1) * (Z - 1);
    double total = 0.0;
#define LNK(x, y, z) links[((int64_t)(x) * Y + (y)) * Z + (z)]
    for (int x = 0; x < X - 1; ++x)
        for (int y = 0; y < Y - 1; ++y)
            for (int z = 0; z < Z - 1; ++z) {
                const int32_t l000 = LNK(x, y, z);
                if (ignore_edge && l000 == 0) continue;
                const int32_t l100 = LNK(x + 1, y, z), l010 = LNK(x, y + 1, z), l001 = LNK(x, y, z + 1);
                for (int idx = start_dim; idx < end_dim; ++idx) {
                    const float v000 = l000 >= 0 ? data[(int64_t)l000 * ddim + idx] : 0.f;
                    const float nullv = ignore_edge ? v000 : 0.f;
                    const float v100 = l100 >= 0 ? data[(int64_t)l100 * ddim + idx] : nullv;
                    const float v010 = l010 >= 0 ? data[(int64_t)l010 * ddim + idx] : nullv;
                    const float v001 = l001 >= 0 ? data[(int64_t)l001 * ddim + idx] : nullv;
                    const float dx = (v100 - v000) * sc[0];
                    const float dy = (v010 - v000) * sc[1];
                    const float dz = (v001 - v000) * sc[2];
                    total += (double)sqrtf(1e-5f + dx * dx + dy * dy + dz * dz);
                }
            }
    return (float)(total * (double)(1.f / (float)nl));
}

void og_tv_grad(const int32_t* links, int32_t X, int32_t Y, int32_t Z, const float* data,
                int32_t ddim, int32_t start_dim, int32_t end_dim, float scale,
                int32_t ignore_edge, float* grad) {
    const float sc[3] = {(float)X * (1.f / 256.f), (float)Y * (1.f / 256.f), (float)Z * (1.f / 256.f)};
    const int nl = (X - 1) * (Y - 1) * (Z - 1);
    const float s = scale / (float)nl;
    for (int x = 0; x < X - 1; ++x)
        for (int y = 0; y < Y - 1; ++y)
            for (int z = 0; z < Z - 1; ++z) {
                const int32_t l000 = LNK(x, y, z);
                if (ignore_edge && l000 == 0) continue;
                const int32_t l100 = LNK(x + 1, y, z), l010 = LNK(x, y + 1, z), l001 = LNK(x, y, z + 1);
                for (int idx = start_dim; idx < end_dim; ++idx) {
                    float v000 = 0.f, v100 = 0.f, v010 = 0.f, v001 = 0.f;
                    if (l000 >= 0) v000 = data[(int64_t)l000 * ddim + idx];
                    if (l100 >= 0) v100 = data[(int64_t)l100 * ddim + idx]; else if (ignore_edge) v100 = v000;
                    if (l010 >= 0) v010 = data[(int64_t)l010 * ddim + idx]; else if (ignore_edge) v010 = v000;
                    if (l001 >= 0) v001 = data[(int64_t)l001 * ddim + idx]; else if (ignore_edge) v001 = v000;
                    float dx = v100 - v000, dy = v010 - v000, dz = v001 - v000;
                    const float idelta = s * (float)(1.0 / sqrt((double)(1e-9f + dx * dx + dy * dy + dz * dz)));
                    dx *= sc[0]; dy *= sc[1]; dz *= sc[2];
                    if (dx != 0.f && l100 >= 0) grad[(int64_t)l100 * ddim + idx] += dx * idelta;
                    if (dy != 0.f && l010 >= 0) grad[(int64_t)l010 * ddim + idx] += dy * idelta;
                    if (dz != 0.f && l001 >= 0) grad[(int64_t)l001 * ddim + idx] += dz * idelta;
                    if (l000 >= 0) grad[(int64_t)l000 * ddim + idx] += -(dx + dy + dz) * idelta;
                }
            }
}

typedef struct {
    const int32_t* links; int32_t X, Y, Z; const float* data; int32_t ddim;
    const int32_t* rand_cells; uint8_t* mask; int32_t start_dim, end_dim; float s;
    int32_t ignore_edge, ignore_last_z; float* grad;
} tvs_ctx_t;

static void tvs_range(int64_t begin, int64_t end, void* vctx) {
    const tvs_ctx_t* c = (const tvs_ctx_t*)vctx;
    const int32_t* links = c->links;
    const int32_t X = c->X, Y = c->Y, Z = c->Z, ddim = c->ddim;
    const float* data = c->data;
    float* grad = c->grad;
    uint8_t* mask = c->mask;
    const int32_t ignore_edge = c->ignore_edge, ignore_last_z = c->ignore_last_z;
    const float s = c->s;
    const float sc[3] = {(float)X * (1.f / 256.f), (float)Y * (1.f / 256.f), (float)Z * (1.f / 256.f)};
    for (int64_t i = begin; i < end; ++i) {
        const int xyz = c->rand_cells[i];
        const int z = xyz % Z;
        const int xy = xyz / Z;
        const int y = xy % Y;
        const int x = xy / Y;
        const int32_t* lp = &LNK(x, y, z);
        if (ignore_edge && *lp == 0) continue;
        const int64_t offx = (int64_t)Y * Z, offy = Z;
        const int32_t lnk000 = lp[0];
        const int32_t lnk001 = ((z + 1 < Z) && (!ignore_last_z || z != Z - 2)) ? lp[1] : 0;
        const int32_t lnk010 = y + 1 < Y ? lp[offy] : 0;
        const int32_t lnk100 = x + 1 < X ? lp[offx] : 0;
        if (ignore_last_z && z == Z - 2) continue;
        for (int idx = c->start_dim; idx < c->end_dim; ++idx) {
            const float v000 = lnk000 >= 0 ? data[(int64_t)lnk000 * ddim + idx] : 0.f;
            const float nullv = ignore_edge ? v000 : 0.f;
            const float v001 = lnk001 >= 0 ? data[(int64_t)lnk001 * ddim + idx] : nullv;
            const float v010 = lnk010 >= 0 ? data[(int64_t)lnk010 * ddim + idx] : nullv;
            const float v100 = lnk100 >= 0 ? data[(int64_t)lnk100 * ddim + idx] : nullv;
            float dx = v100 - v000, dy = v010 - v000, dz = v001 - v000;
            const float idelta = s * (float)(1.0 / sqrt((double)(1e-9f + dx * dx + dy * dy + dz * dz)));
            dx *= sc[0]; dy *= sc[1]; dz *= sc[2];
            const float sm = -(dx + dy + dz);
#define MAYBE_ADD_SET(lnk, val)                                          \
    if ((lnk) >= 0 && (val) != 0.f) {                                    \
        add_float(&grad[(int64_t)(lnk) * ddim + idx], (val) * idelta);   \
        if (mask) mask[(lnk)] = 1;                                       \
    }
            MAYBE_ADD_SET(lnk000, sm);
            MAYBE_ADD_SET(lnk001, dz);
            MAYBE_ADD_SET(lnk010, dy);
            MAYBE_ADD_SET(lnk100, dx);
#undef MAYBE_ADD_SET
        }
    }
}

/* tv_grad_sparse_kernel, loss_kernel.cu:178-246; cells are independent -> parallel_for */
void og_tv_grad_sparse(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                       const float* data, int32_t ddim, const int32_t* rand_cells,
                       int64_t n_cells, uint8_t* mask, int32_t start_dim, int32_t end_dim,
                       float scale, int32_t ignore_edge, int32_t ignore_last_z, float* grad) {
    tvs_ctx_t c = {links, X, Y, Z, data, ddim, rand_cells, mask, start_dim, end_dim,
                   scale / (float)(int)n_cells, ignore_edge, ignore_last_z, grad};
    parallel_for(n_cells, tvs_range, &c);
}
#undef LNK

void og_msi_tv_grad_sparse(const int32_t* bg_links, int32_t reso, const float* msi,
                           int32_t nlayers, int32_t nchan, const int32_t* rand_cells,
                           int64_t n_cells, uint8_t* mask, float scale, float scale_last,
                           float* grad) {
    const int LX = 2 * reso, LY = reso;
    const float s0 = scale / (float)(int)n_cells, s1 = scale_last / (float)(int)n_cells;
#define MSI(l, z, c) msi[(((int64_t)(l) * nlayers) + (z)) * nchan + (c)]
    for (int64_t i = 0; i < n_cells; ++i) {
        const int msi_idx = rand_cells[i];
        const int z = msi_idx % nlayers;
        const int tmp = msi_idx / nlayers;
        const int y = tmp % LY;
        const int x = tmp / LY;
        const int nx = (x == LX - 1) ? 0 : x + 1;
        const int ny = (y == LY - 1) ? 0 : y + 1;
        const int lnk00 = bg_links[x * LY + y];
        const int lnk01 = bg_links[x * LY + ny];
        const int lnk10 = bg_links[nx * LY + y];
        for (int ch = 0; ch < nchan; ++ch) {
            const float v00 = lnk00 >= 0 ? MSI(lnk00, z, ch) : 0.f;
            const float v_nxl = (lnk00 >= 0 && z + 1 < nlayers) ? MSI(lnk00, z + 1, ch)
                                                                 : ((ch == nchan - 1) ? 0.f : v00);
            const float v01 = lnk01 >= 0 ? MSI(lnk01, z, ch) : 0.f;
            const float v10 = lnk10 >= 0 ? MSI(lnk10, z, ch) : 0.f;
            const float sc = (ch == nchan - 1) ? s1 : s0;
            float dx = v10 - v00, dy = v01 - v00, dz = v_nxl - v00;
            const float idelta = sc * (float)(1.0 / sqrt((double)(1e-9f + dx * dx + dy * dy + dz * dz)));
            dx *= (float)LX * (1.f / 256.f);
            dy *= (float)LY * (1.f / 256.f);
            dz *= (float)nlayers * (1.f / 256.f);
            const float sm = -(dx + dy + dz);
#define MAYBE_ADD_SET(link, zz, val)                                              \
    if ((link) >= 0 && (val) != 0.f) {                                            \
        grad[(((int64_t)(link) * nlayers) + (zz)) * nchan + ch] += (val) * idelta; \
        if (mask) mask[(int64_t)(link) * nlayers + (zz)] = 1;                     \
    }
            MAYBE_ADD_SET(lnk00, z, sm);
            if (z + 1 < nlayers) { MAYBE_ADD_SET(lnk00, z + 1, dz); }
            MAYBE_ADD_SET(lnk01, z, dy);
            MAYBE_ADD_SET(lnk10, z, dx);
#undef MAYBE_ADD_SET
        }
    }
#undef MSI
}

/* lumisphere_tv_grad_sparse, loss_kernel.cu:335-434, 643-678.  Cells whose index decodes (with
 * Y-1, Z-1, as upstream) to x >= X-1 are skipped: the reference reads past `links` there. */
void og_lumisphere_tv_grad_sparse(const int32_t* links, int32_t X, int32_t Y, int32_t Z,
                                  const float* sh, int32_t basis_dim, const int32_t* rand_cells,
                                  int64_t n_cells, const float* basis_fn, const float* basis_fn_u,
                                  float scale, float dir_factor, uint8_t* mask, float* grad) {
    const int B = basis_dim, D = 3 * B;
    const int64_t offx = (int64_t)Y * Z;
    const float sc[3] = {(float)X / 256.f, (float)Y / 256.f, (float)Z / 256.f};
    const float scl = scale / (float)n_cells;
    for (int64_t i = 0; i < n_cells; ++i) {
        const int xyz = rand_cells[i];
        const int z = xyz % (Z - 1), xy = xyz / (Z - 1);
        const int y = xy % (Y - 1), x = xy / (Y - 1);
        if (xyz < 0 || x >= X - 1) continue;
        const int32_t* lp = links + (x * offx + (int64_t)y * Z + z);
        if (lp[0] == 0) continue;
        const int32_t L[4] = {lp[0], lp[1], lp[Z], lp[offx]};
        float s[3][5]; /* per colour: sums of v000a, v001a, v010a, v100a, v000u */
        for (int c = 0; c < 3; ++c) {
            for (int k = 0; k < 5; ++k) s[c][k] = 0.f;
            for (int k = 0; k < B; ++k) {
                const int idx = c * B + k;
                const float v000 = L[0] >= 0 ? sh[(int64_t)L[0] * D + idx] : 0.f;
                const float v001 = L[1] >= 0 ? sh[(int64_t)L[1] * D + idx] : v000;
                const float v010 = L[2] >= 0 ? sh[(int64_t)L[2] * D + idx] : v000;
                const float v100 = L[3] >= 0 ? sh[(int64_t)L[3] * D + idx] : v000;
                s[c][0] += v000 * basis_fn[k];
                s[c][1] += v001 * basis_fn[k];
                s[c][2] += v010 * basis_fn[k];
                s[c][3] += v100 * basis_fn[k];
                s[c][4] += v000 * basis_fn_u[k];
            }
        }
        for (int c = 0; c < 3; ++c) {
            float dx = (s[c][3] - s[c][0]) * sc[0];
            float dy = (s[c][2] - s[c][0]) * sc[1];
            float dz = (s[c][1] - s[c][0]) * sc[2];
            float du = (s[c][4] - s[c][0]) * dir_factor;
            const float idelta = scl * (1.f / sqrtf(1e-9f + dx * dx + dy * dy + dz * dz + du * du));
            dx *= sc[0]; dy *= sc[1]; dz *= sc[2]; du *= dir_factor;
            for (int k = 0; k < B; ++k) {
                const int idx = c * B + k;
                const float sv = basis_fn[k];
                const float val[4] = {-dx * sv - dy * sv - dz * sv + du * (basis_fn_u[k] - sv),
                                      dz * sv, dy * sv, dx * sv};
                for (int q = 0; q < 4; ++q)
                    if (L[q] >= 0 && val[q] != 0.f) {
                        grad[(int64_t)L[q] * D + idx] += val[q] * idelta;
                        if (mask) mask[L[q]] = 1;
                    }
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * optimisers: optim_kernel.cu:15-25 (rmsprop_once), :97-103 (sgd_once), dispatch :154-267
 * ------------------------------------------------------------------------------------- */
static inline void rmsprop_once(float* d, float* r, float* gp, float beta, float lr, float eps,
                                float minval) {
    const float g = *gp;
    float rms = *r;
    rms = rms == 0.f ? g * g : lerpf(g * g, rms, beta);
    *r = rms;
    *d = fmaxf(*d - lr * g / (sqrtf(rms) + eps), minval);
    *gp = 0.f;
}

typedef struct {
    float* data; float* rms; float* grad; int32_t C; int32_t kind; const uint8_t* mask;
    const int64_t* index; float beta, lr, eps, minval, lr_last;
} opt_ctx_t;

static void rmsprop_range(int64_t begin, int64_t end, void* vctx) {
    const opt_ctx_t* c = (const opt_ctx_t*)vctx;
    for (int64_t k = begin; k < end; ++k) {
        if (c->kind == 1 && !c->mask[k]) continue;
        const int64_t r = c->kind == 2 ? c->index[k] : k;
        for (int ch = 0; ch < c->C; ++ch) {
            const int64_t off = r * c->C + ch;
            rmsprop_once(c->data + off, c->rms + off, c->grad + off, c->beta,
                         ch == c->C - 1 ? c->lr_last : c->lr, c->eps, c->minval);
        }
    }
}

void og_rmsprop_step(float* data, float* rms, float* grad, int64_t rows, int32_t C,
                     int32_t indexer_kind, const uint8_t* mask, const int64_t* index,
                     int64_t n_index, float beta, float lr, float eps, float minval,
                     float lr_last) {
    if (lr_last < 0.f) lr_last = lr;
    opt_ctx_t c = {data, rms, grad, C, indexer_kind, mask, index, beta, lr, eps, minval, lr_last};
    parallel_for(indexer_kind == 2 ? n_index : rows, rmsprop_range, &c);
}

void og_sgd_step(float* data, float* grad, int64_t rows, int32_t C, int32_t indexer_kind,
                 const uint8_t* mask, const int64_t* index, int64_t n_index, float lr,
                 float lr_last) {
    if (lr_last < 0.f) lr_last = lr;
    if (indexer_kind == 2) {
        for (int64_t k = 0; k < n_index; ++k)
            for (int c = 0; c < C; ++c) {
                const int64_t off = index[k] * C + c;
                data[off] = fmaf(-(c == C - 1 ? lr_last : lr), grad[off], data[off]);
                grad[off] = 0.f;
            }
        return;
    }
    for (int64_t r = 0; r < rows; ++r) {
        if (indexer_kind == 1 && !mask[r]) continue;
        for (int c = 0; c < C; ++c) {
            const int64_t off = r * C + c;
            data[off] = fmaf(-(c == C - 1 ? lr_last : lr), grad[off], data[off]);
            grad[off] = 0.f;
        }
    }
}

/* Thread control for the bench's cpu_baseline leg (1 thread = the scalar port). */
void og_set_num_threads(int n) { g_threads = n > 0 ? n : 1; }
void og_set_accum_float(int on) { g_accum_float = on ? 1 : 0; }
int og_get_max_threads(void) { return g_threads; }
