/* rt_oracle.c -- CPU restatement of kernel_bvh.cl (reference hot path + the
 * per-pixel driver around it), in plain C99.
 *
 * TEST INFRASTRUCTURE ONLY (see rt_oracle.h). Compile with
 *   gcc -std=c99 -O2 -ffp-contract=off -fno-fast-math
 * so every + - * / is one IEEE fp32 rounding, as the conventions in
 * oracle/ref_build/cl_shim.hpp require. Each function cites the reference
 * lines it follows (paths relative to /root/reference).
 */
#define _GNU_SOURCE
#include "rt_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct { float x, y, z; } V3;

static inline V3 v3(float x, float y, float z) { V3 r = { x, y, z }; return r; }
static inline V3 ld(const OrVec* p) { return v3(p->x, p->y, p->z); }
static inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 sdiv(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline V3 normalize(V3 a) { return sdiv(a, sqrtf(dot(a, a))); }
static inline float fmax_cl(float x, float y) { return x < y ? y : x; }   /* OpenCL max() wording */
static inline float fmin_cl(float x, float y) { return y < x ? y : x; }   /* OpenCL min() wording */

/* ---- kernel_bvh.cl:10-16, 42-55 ------------------------------------------------ */
typedef struct { V3 origin, dir, inv; int sign[3]; } Ray;

static Ray init_ray(V3 origin, V3 dir) {
    Ray r;
    dir = normalize(dir);
    r.origin = origin;
    r.dir = dir;
    r.inv = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    r.sign[0] = r.inv.x < 0;
    r.sign[1] = r.inv.y < 0;
    r.sign[2] = r.inv.z < 0;
    return r;
}

/* ---- kernel_bvh.cl:18-27 ------------------------------------------------------- */
typedef struct {
    int hit;
    float t, u, v;
    V3 pos, uv, normal;
    const OrTriangle* object;
} Isect;

/* ---- kernel_bvh.cl:98-153 RayTriangle ------------------------------------------- *
 * The for(i<2) wrapper executes identical work twice with no state change, and
 * CULL_BACKFACE is false, so one pass with the literal reject condition is the
 * same function. Accept is `t < isect->t` only: negative t is accepted.        */
static void ray_triangle(const Ray* r, const OrTriangle* tri, Isect* is) {
    const float HIT_EPSILON = 1.0e-8f;
    V3 t1 = ld(&tri->v1.position), t2 = ld(&tri->v2.position), t3 = ld(&tri->v3.position);
    V3 e1 = sub(t2, t1);
    V3 e2 = sub(t3, t1);
    V3 pvec = cross(r->dir, e2);
    float det = dot(e1, pvec);
    if (det < HIT_EPSILON || -det > HIT_EPSILON) return;           /* :116 */
    float inv_det = 1.0f / det;
    V3 tvec = sub(r->origin, t1);
    float u = dot(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return;                              /* :125 */
    V3 qvec = cross(tvec, e1);
    float v = dot(r->dir, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return;                          /* :132 */
    float t = dot(e2, qvec) * inv_det;
    if (t < is->t) {                                               /* :140 */
        float w = 1.0f - u - v;
        is->hit = 1;
        is->t = t;
        is->u = u;
        is->v = v;
        is->pos = add(r->origin, scale(r->dir, t));
        is->object = tri;
        is->normal = normalize(add(add(scale(ld(&tri->v2.normal), u), scale(ld(&tri->v3.normal), v)),
                                   scale(ld(&tri->v1.normal), w)));
        is->uv = add(add(scale(ld(&tri->v2.uv), u), scale(ld(&tri->v3.uv), v)), scale(ld(&tri->v1.uv), w));
    }
}

/* ---- kernel_bvh.cl:156-169 RayBounds -------------------------------------------- */
static int ray_bounds(const OrNode* n, const Ray* ray, float t) {
    const OrVec* pos[2] = { &n->bmin, &n->bmax };
    float t0 = fmax_cl(0.0f, (pos[ray->sign[0]]->x - ray->origin.x) * ray->inv.x);
    float t1 = fmin_cl(t, (pos[1 - ray->sign[0]]->x - ray->origin.x) * ray->inv.x);
    t0 = fmax_cl(t0, (pos[ray->sign[1]]->y - ray->origin.y) * ray->inv.y);
    t1 = fmin_cl(t1, (pos[1 - ray->sign[1]]->y - ray->origin.y) * ray->inv.y);
    t0 = fmax_cl(t0, (pos[ray->sign[2]]->z - ray->origin.z) * ray->inv.z);
    t1 = fmin_cl(t1, (pos[1 - ray->sign[2]]->z - ray->origin.z) * ray->inv.z);
    return t1 >= t0;
}

/* ---- kernel_bvh.cl:171-219 Intersect -------------------------------------------- *
 * The reference's stack is int[64], unchecked; the oracle grows it on demand so a
 * degenerate tree cannot corrupt memory (results are identical while depth < 64). */
static Isect intersect(const Ray* ray, const OrTriangle* tris, const OrNode* nodes, float tmax,
                       int stop_at_first, OrCounters* c) {
    Isect is;
    memset(&is, 0, sizeof(is));
    is.t = tmax;
    int cap = 64, top = 0, cur = 0;
    int local[64];
    int* stack = local;
    for (;;) {
        const OrNode* node = &nodes[cur];
        if (c) c->nodes_visited++;
        if (ray_bounds(node, ray, is.t)) {
            if (node->nPrimitives > 0) {
                if (c) c->leaves_entered++;
                for (int i = 0; i < node->nPrimitives; ++i) {
                    if (c) c->tris_tested++;
                    ray_triangle(ray, &tris[node->offset + i], &is);
                }
                if (stop_at_first && is.hit) break;
                if (top == 0) break;
                cur = stack[--top];
            } else {
                if (top == cap) {
                    int* bigger = (int*)malloc(sizeof(int) * cap * 2);
                    memcpy(bigger, stack, sizeof(int) * cap);
                    if (stack != local) free(stack);
                    stack = bigger;
                    cap *= 2;
                }
                if (ray->sign[node->axis]) {                       /* :200-203 */
                    stack[top++] = cur + 1;
                    cur = (int)node->offset;
                } else {                                           /* :204-207 */
                    stack[top++] = (int)node->offset;
                    cur = cur + 1;
                }
            }
        } else {
            if (top == 0) break;
            cur = stack[--top];
        }
    }
    if (stack != local) free(stack);
    return is;
}

/* ---- kernel_bvh.cl:57-71 RNG ----------------------------------------------------- */
static inline uint32_t hash_u32(uint32_t x) { return 1103515245u * x + 12345u; }
static inline uint32_t hash_step(uint32_t* x) {
    *x ^= *x >> 16; *x *= 0x7feb352dU; *x ^= *x >> 15; *x *= 0x846ca68bU; *x ^= *x >> 16;
    return *x;
}
static inline float rnd(uint32_t* seed) { return (float)hash_step(seed) / (float)0xffffffffU; }

#define K_TWO_PI 6.28318530718f
#define K_INV_PI 0.31830988618f

/* ---- kernel_bvh.cl:79-90 --------------------------------------------------------- */
static V3 tangent_frame_sample(V3 n, float phi, float sin_theta, float cos_term) {
    V3 axis = fabsf(n.x) > 0.001f ? v3(0.0f, 1.0f, 0.0f) : v3(1.0f, 0.0f, 0.0f);
    V3 t = normalize(cross(axis, n));
    V3 s = cross(n, t);
    return normalize(add(add(scale(scale(s, cosf(phi)), sin_theta), scale(scale(t, sinf(phi)), sin_theta)),
                         scale(n, cos_term)));
}
static V3 sample_hemisphere_cosine(V3 n, uint32_t* seed) {
    float phi = K_TWO_PI * rnd(seed);
    float sin2 = rnd(seed);
    float sin_theta = sqrtf(sin2);
    return tangent_frame_sample(n, phi, sin_theta, sqrtf(1.0f - sin2));
}
/* ---- kernel_bvh.cl:221-239 ------------------------------------------------------- */
static float distribution_ggx(float cos_theta, float alpha) {
    float a2 = alpha * alpha;
    return a2 * K_INV_PI / powf(cos_theta * cos_theta * (a2 - 1.0f) + 1.0f, 2.0f);
}
static V3 sample_ggx(V3 n, float alpha, float* cos_theta, uint32_t* seed) {
    float phi = K_TWO_PI * rnd(seed);
    (void)rnd(seed);                                   /* `xi`: drawn, never used (:230) */
    *cos_theta = powf(rnd(seed), 1.0f / (alpha + 1.0f));
    float sin_theta = sqrtf(fmax_cl(0.0f, 1.0f - (*cos_theta) * (*cos_theta)));
    return tangent_frame_sample(n, phi, sin_theta, *cos_theta);
}
/* ---- kernel_bvh.cl:264-302 (G and F are computed by the reference but never
 * reach the returned colour or pdf, so they are omitted here) ---------------------- */
static V3 sample_brdf(V3 wo, V3* wi, float* pdf, V3 normal, const OrMaterial* m, uint32_t* seed) {
    if (rnd(seed) > 0.5f) {                            /* specular, :271-292 */
        float cos_theta = 1;
        float alpha = 2.0f / powf(m->roughness, 2.0f) - 2.0f;
        V3 wh = sample_ggx(normal, alpha, &cos_theta, seed);
        *wi = add(neg(wo), scale(wh, 2.0f * dot(wo, wh)));          /* reflect(), :74-77 */
        if (dot(*wi, normal) * dot(wo, normal) < 0.000001f) return v3(0.0f, 0.0f, 0.0f);
        float D = distribution_ggx(cos_theta, alpha);
        *pdf = D * cos_theta / (4.0f * fmax_cl(dot(wo, wh), 0.0f));
        float k = D / ((4.0f * fmax_cl(dot(*wi, normal), 0.0f) * fmax_cl(dot(wo, normal), 0.0f)) + 0.001f);
        return v3(k * m->specular.x, k * m->specular.y, k * m->specular.z);
    }
    *wi = sample_hemisphere_cosine(normal, seed);      /* diffuse, :264-269 */
    *pdf = dot(*wi, normal) * K_INV_PI;
    return scale(ld(&m->diffuse), K_INV_PI);
}
/* ---- kernel_bvh.cl:304-347 ------------------------------------------------------- */
static float light_pixel(const Ray* ray, int light_type, const Isect* is) {
    V3 light_pos = v3(0.0f, -10.0f, 16.0f);
    V3 light_dir = v3(-0.5f, 0.4f, -0.1f);
    float intensity = 1.0f, ndotl = 1.0f, attn = 1.0f;
    if (light_type <= 0) {
        ndotl = fmax_cl(dot(is->normal, neg(light_dir)), 0.0f);
    } else if (light_type == 1) {
        intensity = 16.0f;
        float falloff = 0.8f;
        V3 X = add(ray->origin, scale(ray->dir, is->t));
        V3 L = sub(light_pos, X);
        ndotl = fmax_cl(dot(is->normal, L), 0.0f);
        V3 eye = sub(L, X);
        float d = sqrtf(dot(eye, eye));
        attn = (float)(1.0 / (double)(falloff * (d * d)));         /* `1.0 /` is a double literal (:335) */
    } else {
        V3 X = add(ray->origin, scale(ray->dir, is->t));
        V3 L = sub(light_pos, X);
        ndotl = fmax_cl(dot(is->normal, L), 0.0f);
    }
    return attn * intensity * ndotl;
}
/* ---- kernel_bvh.cl:349-384 ------------------------------------------------------- */
static V3 render_path(Ray* ray, const OrTriangle* tris, const OrNode* nodes, const OrMaterial* mats,
                      uint32_t bounces, int light_type, float sky, uint32_t* seed) {
    V3 radiance = v3(0, 0, 0), beta = v3(1, 1, 1);
    for (int i = 0; (uint32_t)i < bounces; ++i) {
        Isect is = intersect(ray, tris, nodes, 100000.0f, 0, NULL);
        if (!is.hit) {
            radiance = add(radiance, mul(beta, scale(v3(0.5f, 0.5f, 0.5f), sky)));
            break;
        }
        const OrMaterial* m = &mats[is.object->mtlIndex];
        radiance = add(radiance, scale(mul(beta, ld(&m->emission)), 50.0f));
        V3 wi = v3(0, 0, 0), wo = neg(ray->dir);
        float pdf = 0.0f;
        V3 f = sample_brdf(wo, &wi, &pdf, is.normal, m, seed);
        if (pdf <= 0.0f || pdf != pdf) break;
        beta = mul(beta, sdiv(scale(f, dot(wi, is.normal)), pdf));
        float lp = light_pixel(ray, light_type, &is);
        radiance = add(radiance, mul(mul(v3(lp, lp, lp), ld(&m->diffuse)), beta));
        *ray = init_ray(add(is.pos, scale(wi, 0.01f)), wi);
    }
    return v3(fmax_cl(radiance.x, 0.0f), fmax_cl(radiance.y, 0.0f), fmax_cl(radiance.z, 0.0f));
}
/* ---- kernel_bvh.cl:386-403 ------------------------------------------------------- */
static Ray create_ray(uint32_t gid, uint32_t width, uint32_t height, V3 cam_pos, V3 cam_front, V3 cam_up,
                      uint32_t* seed) {
    float inv_w = 1.0f / (float)width, inv_h = 1.0f / (float)height;
    float aspect = (float)width / (float)height;
    float angle = tanf(0.5f * (45.0f * 3.1415f / 180.0f));
    float x = (float)(gid % width) + rnd(seed) - 0.5f;
    float y = (float)(gid / width) + rnd(seed) - 0.5f;
    x = (2.0f * ((x + 0.5f) * inv_w) - 1) * angle * aspect;
    y = -(1.0f - 2.0f * ((y + 0.5f) * inv_h)) * angle;
    V3 dir = normalize(add(add(scale(cross(cam_front, cam_up), x), scale(cam_up, y)), cam_front));
    return init_ray(cam_pos, dir);
}

/* ---- threading helper ------------------------------------------------------------ */
typedef void (*range_fn)(uint64_t lo, uint64_t hi, void* ctx, int worker);
typedef struct { range_fn fn; void* ctx; uint64_t lo, hi; int worker; } Job;
static void* job_main(void* p) { Job* j = (Job*)p; j->fn(j->lo, j->hi, j->ctx, j->worker); return NULL; }

int oracle_hardware_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
static int resolve_threads(int threads, uint64_t n) {
    if (threads <= 0) threads = oracle_hardware_threads();
    if (threads > 1024) threads = 1024;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    return threads;
}
static void parallel_ranges(uint64_t lo, uint64_t hi, int threads, range_fn fn, void* ctx) {
    uint64_t n = hi > lo ? hi - lo : 0;
    threads = resolve_threads(threads, n);
    if (threads <= 1) { fn(lo, hi, ctx, 0); return; }
    pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    Job* jobs = (Job*)malloc(sizeof(Job) * threads);
    /* many small interleaved chunks would balance better, but contiguous blocks
       keep the code obvious; callers shuffle rays if they need balance */
    for (int i = 0; i < threads; ++i) {
        jobs[i].fn = fn; jobs[i].ctx = ctx; jobs[i].worker = i;
        jobs[i].lo = lo + n * (uint64_t)i / (uint64_t)threads;
        jobs[i].hi = lo + n * (uint64_t)(i + 1) / (uint64_t)threads;
        pthread_create(&tid[i], NULL, job_main, &jobs[i]);
    }
    for (int i = 0; i < threads; ++i) pthread_join(tid[i], NULL);
    free(tid);
    free(jobs);
}

/* ---- public entry points ---------------------------------------------------------- */
typedef struct {
    const OrTriangle* tris; const OrNode* nodes; const OrRay* rays;
    OrHit* hits; OrHitAttr* attr; uint32_t* occ; OrCounters* per_worker;
} TraceCtx;

static void closest_range(uint64_t lo, uint64_t hi, void* p, int worker) {
    TraceCtx* c = (TraceCtx*)p;
    OrCounters* cnt = c->per_worker ? &c->per_worker[worker] : NULL;
    for (uint64_t i = lo; i < hi; ++i) {
        const OrRay* q = &c->rays[i];
        Ray r = init_ray(v3(q->ox, q->oy, q->oz), v3(q->dx, q->dy, q->dz));
        Isect is = intersect(&r, c->tris, c->nodes, q->tmax, 0, cnt);
        OrHit h;
        h.t = is.t; h.u = is.hit ? is.u : 0.0f; h.v = is.hit ? is.v : 0.0f;
        h.tri = is.hit ? (uint32_t)(is.object - c->tris) : 0xFFFFFFFFu;
        c->hits[i] = h;
        if (c->attr) {
            OrHitAttr a;
            memset(&a, 0, sizeof(a));
            if (is.hit) {
                a.pos[0] = is.pos.x; a.pos[1] = is.pos.y; a.pos[2] = is.pos.z;
                a.normal[0] = is.normal.x; a.normal[1] = is.normal.y; a.normal[2] = is.normal.z;
                a.uv[0] = is.uv.x; a.uv[1] = is.uv.y;
            }
            c->attr[i] = a;
        }
    }
}
static void any_range(uint64_t lo, uint64_t hi, void* p, int worker) {
    (void)worker;
    TraceCtx* c = (TraceCtx*)p;
    for (uint64_t i = lo; i < hi; ++i) {
        const OrRay* q = &c->rays[i];
        Ray r = init_ray(v3(q->ox, q->oy, q->oz), v3(q->dx, q->dy, q->dz));
        c->occ[i] = (uint32_t)intersect(&r, c->tris, c->nodes, q->tmax, 1, NULL).hit;
    }
}

void oracle_trace_closest(const OrTriangle* tris, const OrNode* nodes, const OrRay* rays, uint64_t n,
                          OrHit* hits, OrHitAttr* attr, OrCounters* counters, int threads) {
    TraceCtx c = { tris, nodes, rays, hits, attr, NULL, NULL };
    int nt = resolve_threads(threads, n);
    if (counters) c.per_worker = (OrCounters*)calloc((size_t)nt, sizeof(OrCounters));
    parallel_ranges(0, n, nt, closest_range, &c);
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (int i = 0; i < nt; ++i) {
            counters->nodes_visited += c.per_worker[i].nodes_visited;
            counters->leaves_entered += c.per_worker[i].leaves_entered;
            counters->tris_tested += c.per_worker[i].tris_tested;
        }
        free(c.per_worker);
    }
}

void oracle_trace_any(const OrTriangle* tris, const OrNode* nodes, const OrRay* rays, uint64_t n,
                      uint32_t* occluded, int threads) {
    TraceCtx c = { tris, nodes, rays, NULL, NULL, occluded, NULL };
    parallel_ranges(0, n, threads, any_range, &c);
}

typedef struct {
    const OrTriangle* tris; const OrNode* nodes; const OrMaterial* mats; float* result;
    uint32_t width, height, frame_count; int32_t bounces, light_type; float sky;
    V3 pos, front, up; OrRay* rays; uint64_t gid0;
} RenderCtx;

/* ---- kernel_bvh.cl:415-456 KernelEntry ---------------------------------------------- */
static void render_range(uint64_t lo, uint64_t hi, void* p, int worker) {
    (void)worker;
    RenderCtx* c = (RenderCtx*)p;
    for (uint64_t g = lo; g < hi; ++g) {
        uint32_t gid = (uint32_t)g;
        uint32_t seed = gid + hash_u32(c->frame_count);
        Ray ray = create_ray(gid, c->width, c->height, c->pos, c->front, c->up, &seed);
        V3 rad = render_path(&ray, c->tris, c->nodes, c->mats, (uint32_t)c->bounces, c->light_type, c->sky, &seed);
        float* px = c->result + 4 * g;
        if (c->frame_count == 0) {
            px[0] = powf(rad.x, 0.45454545f); px[1] = powf(rad.y, 0.45454545f); px[2] = powf(rad.z, 0.45454545f);
        } else {
            float fm1 = (float)(c->frame_count - 1), fc = (float)c->frame_count;
            float in[3] = { px[0], px[1], px[2] }, r[3] = { rad.x, rad.y, rad.z };
            for (int k = 0; k < 3; ++k)
                px[k] = powf((powf(in[k], 2.2f) * fm1 + r[k]) / fc, 0.454545f);
        }
        px[3] = 0.0f;
    }
}

void oracle_render(const OrTriangle* tris, const OrNode* nodes, const OrMaterial* mats, float* result,
                   uint32_t width, uint32_t height, uint32_t frame_count,
                   int32_t light_bounces, int32_t light_type, float sky,
                   const float* cam_pos, const float* cam_front, const float* cam_up,
                   uint64_t gid0, uint64_t gid1, int threads) {
    RenderCtx c;
    memset(&c, 0, sizeof(c));
    c.tris = tris; c.nodes = nodes; c.mats = mats; c.result = result;
    c.width = width; c.height = height; c.frame_count = frame_count;
    c.bounces = light_bounces; c.light_type = light_type; c.sky = sky;
    c.pos = v3(cam_pos[0], cam_pos[1], cam_pos[2]);
    c.front = v3(cam_front[0], cam_front[1], cam_front[2]);
    c.up = v3(cam_up[0], cam_up[1], cam_up[2]);
    parallel_ranges(gid0, gid1, threads, render_range, &c);
}

static void camera_range(uint64_t lo, uint64_t hi, void* p, int worker) {
    (void)worker;
    RenderCtx* c = (RenderCtx*)p;
    for (uint64_t g = lo; g < hi; ++g) {
        uint32_t gid = (uint32_t)g;
        uint32_t seed = gid + hash_u32(c->frame_count);
        Ray r = create_ray(gid, c->width, c->height, c->pos, c->front, c->up, &seed);
        OrRay o = { r.origin.x, r.origin.y, r.origin.z, 0.0f, r.dir.x, r.dir.y, r.dir.z, 100000.0f };
        c->rays[g - c->gid0] = o;
    }
}

void oracle_camera_rays(uint32_t width, uint32_t height, uint32_t frame_count,
                        const float* cam_pos, const float* cam_front, const float* cam_up,
                        uint64_t gid0, uint64_t gid1, OrRay* rays) {
    RenderCtx c;
    memset(&c, 0, sizeof(c));
    c.width = width; c.height = height; c.frame_count = frame_count;
    c.pos = v3(cam_pos[0], cam_pos[1], cam_pos[2]);
    c.front = v3(cam_front[0], cam_front[1], cam_front[2]);
    c.up = v3(cam_up[0], cam_up[1], cam_up[2]);
    c.rays = rays; c.gid0 = gid0;
    parallel_ranges(gid0, gid1, 1, camera_range, &c);
}
