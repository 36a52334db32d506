/*
 * maray_oracle.c -- CPU ORACLE for the Maray per-pixel render path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference's CPU algorithm for the hot path
 * (advancedresearch/maray v0.3.8).  It exists so the CUDA path can be checked against the
 * reference's arithmetic; it is NOT part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (maray_b200/, include/) never links, imports or calls anything in oracle/.
 *
 * Why a restatement: the reference is Rust + wasmer; no cargo/rustc exists in this image or on
 * the GPU box, so the reference itself cannot be compiled or run (DESIGN.md "Oracle").
 *
 * Pinning (see tests/test_oracle_golden.py):
 *   - every known-answer assertion of the reference's `it_works` test (src/lib.rs:1241-1285);
 *   - data/chess.maray parses to exact EOF in the legacy layout;
 *   - the render of data/chess.maray against images/chess.png (approximate golden).
 *   sin/exp/ln go to this host's glibc, exactly what the Rust binary would call here;
 *   beyond cos(0)==1.0 the reference pins nothing at that boundary ("parity unpinned" there).
 *
 * Each function cites the reference file:line it follows (paths relative to the reference root).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Expr IR: src/lib.rs:101-149 (HEAD variant numbering).                                        */

enum {
    T_ARC = 0, T_X, T_Y, T_TAU, T_E, T_VAR, T_NAT,
    T_NEG, T_ABS, T_RECIP, T_SQRT, T_STEP, T_SIN, T_EXP, T_LN,
    T_ADD, T_MUL, T_MAX, T_MIN, T_LET, T_DECOR, T_APP, T_COUNT
};

typedef struct Expr Expr;
typedef struct Ctx Ctx;

/* Context: src/lib.rs:51-55 -- Vec<(u64, Expr)>. */
struct Ctx {
    uint64_t n;
    uint64_t *ids;
    Expr **defs;
};

struct Expr {
    uint32_t tag;
    uint32_t app_id;   /* App: function id (u32)                    */
    uint64_t n;        /* Var: id, Nat: value                       */
    Expr *a, *b;       /* unary: a; binary/App: a,b; Let: body in a */
    Ctx *ctx;          /* Let only                                  */
    uint64_t hash;     /* structural hash, filled by fix()          */
};

/* Arena: nodes live as long as the scene. */
typedef struct Block { struct Block *next; size_t used, cap; } Block;
typedef struct { Block *head; } Arena;

static void *arena_alloc(Arena *ar, size_t sz) {
    sz = (sz + 15) & ~(size_t)15;
    if (!ar->head || ar->head->used + sz > ar->head->cap) {
        size_t cap = sz > (1u << 20) ? sz : (1u << 20);
        Block *b = (Block *)malloc(sizeof(Block) + cap);
        if (!b) { fprintf(stderr, "maray_oracle: out of memory\n"); abort(); }
        b->next = ar->head; b->used = 0; b->cap = cap; ar->head = b;
    }
    void *p = (char *)(ar->head + 1) + ar->head->used;
    ar->head->used += sz;
    memset(p, 0, sz);
    return p;
}
static void arena_free(Arena *ar) {
    Block *b = ar->head;
    while (b) { Block *n = b->next; free(b); b = n; }
    ar->head = NULL;
}
static Expr *mk(Arena *ar, uint32_t tag) {
    Expr *e = (Expr *)arena_alloc(ar, sizeof(Expr));
    e->tag = tag;
    return e;
}

/* ------------------------------------------------------------------------------------------ */
/* Wire format: src/lib.rs:1227-1235 `open` = bincode 1.3.3 default options (little-endian,
 * fixed-width ints, u32 enum tags, u64 lengths; Box/Arc transparent).  HEAD numbering is the
 * enum order at src/lib.rs:101-149; the legacy numbering (no `Arc` variant, every tag one
 * lower) is what data/chess.maray uses.                           */

typedef struct {
    const uint8_t *p, *end;
    int legacy;
    int err;
    Arena *ar;
    uint32_t depth;
} Rd;

static uint32_t rd_u32(Rd *r) {
    if (r->err || (size_t)(r->end - r->p) < 4) { r->err = 1; return 0; }
    uint32_t v; memcpy(&v, r->p, 4); r->p += 4; return v;
}
static uint64_t rd_u64(Rd *r) {
    if (r->err || (size_t)(r->end - r->p) < 8) { r->err = 1; return 0; }
    uint64_t v; memcpy(&v, r->p, 8); r->p += 8; return v;
}

static Expr *rd_expr(Rd *r);

/* Token: src/token.rs:11-38.  Decor is semantically transparent (src/lib.rs:663) so tokens are
 * parsed only to find where they end. */
static void rd_token(Rd *r) {
    uint32_t t = rd_u32(r);
    if (r->err) return;
    if (t == 0) { (void)rd_expr(r); }
    else if (t == 1) {
        uint64_t len = rd_u64(r);
        if (r->err || (uint64_t)(r->end - r->p) < len) { r->err = 1; return; }
        r->p += len;
    } else if (t > 12) r->err = 1;
}

static Expr *rd_expr(Rd *r) {
    if (r->err) return NULL;
    if (++r->depth > 200000) { r->err = 1; return NULL; }
    uint32_t t = rd_u32(r);
    if (r->err) return NULL;
    if (r->legacy) t += 1;                       /* legacy file has no Arc variant */
    if (t >= T_COUNT || (r->legacy && t == T_ARC)) { r->err = 1; return NULL; }
    Expr *e = mk(r->ar, t);
    switch (t) {
    case T_ARC: {                                /* serde "rc": Arc<Expr> is its inner value */
        e->a = rd_expr(r);
        break;
    }
    case T_X: case T_Y: case T_TAU: case T_E: break;
    case T_VAR: case T_NAT: e->n = rd_u64(r); break;
    case T_NEG: case T_ABS: case T_RECIP: case T_SQRT:
    case T_STEP: case T_SIN: case T_EXP: case T_LN:
        e->a = rd_expr(r); break;
    case T_ADD: case T_MUL: case T_MAX: case T_MIN:
        e->a = rd_expr(r); e->b = rd_expr(r); break;
    case T_LET: {
        uint64_t n = rd_u64(r);
        if (r->err || n > (uint64_t)(r->end - r->p) / 12) { r->err = 1; return NULL; }
        Ctx *c = (Ctx *)arena_alloc(r->ar, sizeof(Ctx));
        c->n = n;
        c->ids = (uint64_t *)arena_alloc(r->ar, (n ? n : 1) * sizeof(uint64_t));
        c->defs = (Expr **)arena_alloc(r->ar, (n ? n : 1) * sizeof(Expr *));
        for (uint64_t i = 0; i < n && !r->err; i++) {
            c->ids[i] = rd_u64(r);
            c->defs[i] = rd_expr(r);
        }
        e->ctx = c;
        e->a = rd_expr(r);
        break;
    }
    case T_DECOR: {
        e->a = rd_expr(r);
        uint64_t n = rd_u64(r);
        for (uint64_t i = 0; i < n && !r->err; i++) rd_token(r);
        break;
    }
    case T_APP:
        e->app_id = rd_u32(r);
        e->a = rd_expr(r); e->b = rd_expr(r); break;
    }
    r->depth--;
    return r->err ? NULL : e;
}

/* ------------------------------------------------------------------------------------------ */
/* var_fixer: src/var_fixer.rs:25-82.  Restated as written, including that sibling references
 * inside a Let's definitions are renamed with the OUTER mapping (src/var_fixer.rs:51-52;
 * SURVEY.md F6).  `ids: HashMap<Expr,u64>` is keyed on structural equality of the fixed
 * expression; here a structural hash + deep compare.                                          */

static uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    return h ^ (h >> 32);
}

static int expr_eq(const Expr *x, const Expr *y) {
    if (x == y) return 1;
    if (x->hash != y->hash || x->tag != y->tag) return 0;
    switch (x->tag) {
    case T_X: case T_Y: case T_TAU: case T_E: return 1;
    case T_VAR: case T_NAT: return x->n == y->n;
    case T_NEG: case T_ABS: case T_RECIP: case T_SQRT:
    case T_STEP: case T_SIN: case T_EXP: case T_LN: case T_ARC: case T_DECOR:
        /* Decor tokens were dropped at parse time; scenes on the render path carry none. */
        return expr_eq(x->a, y->a);
    case T_ADD: case T_MUL: case T_MAX: case T_MIN:
        return expr_eq(x->a, y->a) && expr_eq(x->b, y->b);
    case T_APP:
        return x->app_id == y->app_id && expr_eq(x->a, y->a) && expr_eq(x->b, y->b);
    case T_LET:
        if (x->ctx->n != y->ctx->n) return 0;
        for (uint64_t i = 0; i < x->ctx->n; i++)
            if (x->ctx->ids[i] != y->ctx->ids[i] || !expr_eq(x->ctx->defs[i], y->ctx->defs[i])) return 0;
        return expr_eq(x->a, y->a);
    }
    return 0;
}

typedef struct { Expr **keys; uint64_t *vals; size_t cap, len; } IdMap;
typedef struct { uint64_t old_id, new_id; } Ren;
typedef struct { Ren *v; size_t n; } RenCtx;

typedef struct {
    Arena *ar;
    IdMap ids;           /* VarFixer::ids       src/var_fixer.rs:10 */
    uint64_t var_count;  /* VarFixer::var_count src/var_fixer.rs:12 */
} Fixer;

static void idmap_grow(IdMap *m) {
    size_t ncap = m->cap ? m->cap * 2 : 1024;
    Expr **nk = (Expr **)calloc(ncap, sizeof(Expr *));
    uint64_t *nv = (uint64_t *)calloc(ncap, sizeof(uint64_t));
    for (size_t i = 0; i < m->cap; i++) if (m->keys[i]) {
        size_t j = m->keys[i]->hash & (ncap - 1);
        while (nk[j]) j = (j + 1) & (ncap - 1);
        nk[j] = m->keys[i]; nv[j] = m->vals[i];
    }
    free(m->keys); free(m->vals);
    m->keys = nk; m->vals = nv; m->cap = ncap;
}
static int idmap_get(IdMap *m, Expr *k, uint64_t *out) {
    if (!m->cap) return 0;
    size_t j = k->hash & (m->cap - 1);
    while (m->keys[j]) {
        if (expr_eq(m->keys[j], k)) { *out = m->vals[j]; return 1; }
        j = (j + 1) & (m->cap - 1);
    }
    return 0;
}
static void idmap_put(IdMap *m, Expr *k, uint64_t v) {
    if ((m->len + 1) * 2 > m->cap) idmap_grow(m);
    size_t j = k->hash & (m->cap - 1);
    while (m->keys[j]) j = (j + 1) & (m->cap - 1);
    m->keys[j] = k; m->vals[j] = v; m->len++;
}

/* VarFixer::fix  src/var_fixer.rs:25-70 */
static Expr *fix(Fixer *f, const Expr *e, const RenCtx *ctx) {
    Expr *o;
    switch (e->tag) {
    case T_ARC:                                                   /* :29 Arc is unwrapped */
        return fix(f, e->a, ctx);
    case T_X: case T_Y: case T_TAU: case T_E: case T_NAT:         /* :30 */
        o = mk(f->ar, e->tag); o->n = e->n;
        o->hash = mix(e->tag, e->n);
        return o;
    case T_VAR:                                                   /* :31-36 first match wins */
        o = mk(f->ar, T_VAR); o->n = e->n;
        for (size_t i = 0; i < ctx->n; i++)
            if (e->n == ctx->v[i].old_id) { o->n = ctx->v[i].new_id; break; }
        o->hash = mix(T_VAR, o->n);
        return o;
    case T_NEG: case T_ABS: case T_RECIP: case T_SQRT:
    case T_STEP: case T_SIN: case T_EXP: case T_LN:               /* :37-44 */
        o = mk(f->ar, e->tag); o->a = fix(f, e->a, ctx);
        o->hash = mix(e->tag, o->a->hash);
        return o;
    case T_ADD: case T_MUL: case T_MAX: case T_MIN:               /* :45-48 */
        o = mk(f->ar, e->tag); o->a = fix(f, e->a, ctx); o->b = fix(f, e->b, ctx);
        o->hash = mix(mix(e->tag, o->a->hash), o->b->hash);
        return o;
    case T_LET: {                                                 /* :49-67 */
        Ctx *c = (Ctx *)arena_alloc(f->ar, sizeof(Ctx));
        uint64_t n = e->ctx->n;
        c->n = n;
        c->ids = (uint64_t *)arena_alloc(f->ar, (n ? n : 1) * sizeof(uint64_t));
        c->defs = (Expr **)arena_alloc(f->ar, (n ? n : 1) * sizeof(Expr *));
        RenCtx nctx; nctx.n = 0;
        nctx.v = (Ren *)malloc((n ? n : 1) * sizeof(Ren));
        uint64_t h = mix(T_LET, n);
        for (uint64_t i = 0; i < n; i++) {
            Expr *d = fix(f, e->ctx->defs[i], ctx);               /* :52 outer ctx (F6) */
            uint64_t id;
            if (!idmap_get(&f->ids, d, &id)) {                    /* :53-62 */
                id = f->var_count++;
                idmap_put(&f->ids, d, id);
            }
            nctx.v[nctx.n].old_id = e->ctx->ids[i];
            nctx.v[nctx.n].new_id = id;
            nctx.n++;
            c->ids[i] = id;
            c->defs[i] = d;
            h = mix(mix(h, id), d->hash);
        }
        o = mk(f->ar, T_LET);
        o->ctx = c;
        o->a = fix(f, e->a, &nctx);                               /* :66 */
        o->hash = mix(h, o->a->hash);
        free(nctx.v);
        return o;
    }
    case T_DECOR:                                                 /* :68 */
        o = mk(f->ar, T_DECOR); o->a = fix(f, e->a, ctx);
        o->hash = mix(T_DECOR, o->a->hash);
        return o;
    case T_APP:                                                   /* :69 */
        o = mk(f->ar, T_APP); o->app_id = e->app_id;
        o->a = fix(f, e->a, ctx); o->b = fix(f, e->b, ctx);
        o->hash = mix(mix(mix(T_APP, e->app_id), o->a->hash), o->b->hash);
        return o;
    }
    return NULL;
}

/* ------------------------------------------------------------------------------------------ */
/* Runtime<Textures>: src/lib.rs:72-98, src/textures.rs:9-65.                                   */

typedef struct {
    uint32_t n;
    uint8_t **data;      /* RGB8, row-major, no padding (image::RgbImage) */
    uint32_t *w, *h;
    /* Runtime::functions has 5*n entries (src/textures.rs:54-65). */
} Textures;

#define ALIGN 5u   /* src/textures.rs:14 */

/* `f64 as u32` (Rust saturating cast): NaN -> 0, clamp to [0, u32::MAX], truncate. */
static uint32_t f64_as_u32(double v) {
    if (!(v > 0.0)) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}

/* fun_color_channel src/textures.rs:27-36; fun_image_width :40-43; fun_image_height :47-50 */
static double texture_fn(const Textures *t, uint32_t id, double x, double y) {
    uint32_t img = id / ALIGN, k = id % ALIGN;
    if (k == 3) return (double)t->w[img];
    if (k == 4) return (double)t->h[img];
    uint32_t w = t->w[img], h = t->h[img];
    if (x < 0.0 || y < 0.0) return 0.0;
    uint32_t xi = f64_as_u32(x), yi = f64_as_u32(y);
    if (xi >= w || yi >= h) return 0.0;
    return (double)t->data[img][((size_t)yi * w + xi) * 3 + k];
}

/* ------------------------------------------------------------------------------------------ */
/* Cache: src/cache.rs:6-42 -- FnvHashMap<u64,(f64,bool)>.                                      */

typedef struct { uint64_t key; double val; uint8_t depx, used; } CEnt;
typedef struct {
    CEnt *tab; size_t cap, len;
    CEnt *tmp;           /* scratch for retain() */
} Cache;

static void cache_init(Cache *c) {
    c->cap = 256; c->len = 0;
    c->tab = (CEnt *)calloc(c->cap, sizeof(CEnt));
    c->tmp = (CEnt *)calloc(c->cap, sizeof(CEnt));
}
static void cache_free(Cache *c) { free(c->tab); free(c->tmp); }
static size_t cache_slot(uint64_t k, size_t cap) {
    uint64_t h = 0xcbf29ce484222325ull;           /* FNV-1a over the 8 key bytes, as fnv::FnvHasher */
    for (int i = 0; i < 8; i++) { h ^= (k >> (8 * i)) & 0xff; h *= 0x100000001b3ull; }
    return (size_t)(h ^ (h >> 29)) & (cap - 1);
}
static void cache_put_raw(CEnt *tab, size_t cap, uint64_t k, double v, uint8_t depx) {
    size_t j = cache_slot(k, cap);
    while (tab[j].used) j = (j + 1) & (cap - 1);
    tab[j].key = k; tab[j].val = v; tab[j].depx = depx; tab[j].used = 1;
}
static void cache_insert(Cache *c, uint64_t k, double v, uint8_t depx) {
    if ((c->len + 1) * 2 > c->cap) {
        size_t ncap = c->cap * 2;
        CEnt *nt = (CEnt *)calloc(ncap, sizeof(CEnt));
        for (size_t i = 0; i < c->cap; i++)
            if (c->tab[i].used) cache_put_raw(nt, ncap, c->tab[i].key, c->tab[i].val, c->tab[i].depx);
        free(c->tab); free(c->tmp);
        c->tab = nt; c->cap = ncap;
        c->tmp = (CEnt *)calloc(ncap, sizeof(CEnt));
    }
    /* HashMap::insert overwrites an existing key */
    size_t j = cache_slot(k, c->cap);
    while (c->tab[j].used) {
        if (c->tab[j].key == k) { c->tab[j].val = v; c->tab[j].depx = depx; return; }
        j = (j + 1) & (c->cap - 1);
    }
    c->tab[j].key = k; c->tab[j].val = v; c->tab[j].depx = depx; c->tab[j].used = 1;
    c->len++;
}
static CEnt *cache_get(Cache *c, uint64_t k) {
    size_t j = cache_slot(k, c->cap);
    while (c->tab[j].used) {
        if (c->tab[j].key == k) return &c->tab[j];
        j = (j + 1) & (c->cap - 1);
    }
    return NULL;
}
/* Cache::clear src/cache.rs:15 */
static void cache_clear(Cache *c) { memset(c->tab, 0, c->cap * sizeof(CEnt)); c->len = 0; }
/* Cache::clear_dep_x src/cache.rs:18-20: retain entries that do not depend on x. */
static void cache_clear_dep_x(Cache *c) {
    if (!c->len) return;
    size_t keep = 0, drop = 0;
    for (size_t i = 0; i < c->cap; i++) if (c->tab[i].used) {
        if (c->tab[i].depx) drop++; else c->tmp[keep++] = c->tab[i];
    }
    if (!drop) return;
    memset(c->tab, 0, c->cap * sizeof(CEnt));
    for (size_t i = 0; i < keep; i++) cache_put_raw(c->tab, c->cap, c->tmp[i].key, c->tmp[i].val, c->tmp[i].depx);
    c->len = keep;
}

/* ------------------------------------------------------------------------------------------ */
/* Interpreter: Expr::eval2 src/lib.rs:623-670, Expr::dep_x src/lib.rs:675-706,
 * Cache::val src/cache.rs:23-42.                                                              */

typedef struct {
    const Textures *tex;      /* rt.ctx */
    uint32_t n_functions;     /* rt.functions.len() */
    int fault;                /* set when the Rust code would panic (functions[id] out of range) */
    double *probe;            /* diagnostic (mo_step_margin): smallest relative |step argument| seen, in ULP */
} Rt;

static const Ctx EMPTY_CTX = {0, NULL, NULL};

/* f64::max / f64::min (src/lib.rs:655-658).  Rust lowers these to llvm.maxnum/minnum; on
 * x86-64 that is `select(isnan(a), b, MAXSD(b, a))`, and MAXSD(b, a) = (b > a) ? b : a.  So a
 * NaN operand is ignored and an equal-compare tie (including +0 / -0) returns `a` (self).     */
static inline double rust_max(double a, double b) { return (a != a) ? b : ((b > a) ? b : a); }
static inline double rust_min(double a, double b) { return (a != a) ? b : ((b < a) ? b : a); }

typedef struct { double v; int depx; } ValDep;

static double eval2(const Expr *e, Rt *rt, const double v[2], const Ctx *ctx, Cache *cache);
static void probe_step(const Expr *arg, double s, Rt *rt, const double v[2], const Ctx *ctx, Cache *cache);
static int dep_x(const Expr *e, Rt *rt, const double v[2], const Ctx *ctx, Cache *cache);

/* Cache::val src/cache.rs:23-42 */
static ValDep cache_val(Cache *cache, Rt *rt, const double p[2], uint64_t name, const Ctx *ctx) {
    ValDep r;
    CEnt *hit = cache_get(cache, name);
    if (hit) { r.v = hit->val; r.depx = hit->depx; return r; }
    for (uint64_t i = 0; i < ctx->n; i++) {
        if (ctx->ids[i] == name) {
            double val = eval2(ctx->defs[i], rt, p, ctx, cache);          /* :33 */
            int d = dep_x(ctx->defs[i], rt, p, ctx, cache);               /* :34 second walk */
            cache_insert(cache, name, val, (uint8_t)d);                   /* :35 */
            r.v = val; r.depx = d; return r;
        }
    }
    r.v = NAN; r.depx = 0;                                                /* :40 unbound -> NaN */
    return r;
}

static double eval2(const Expr *e, Rt *rt, const double v[2], const Ctx *ctx, Cache *cache) {
    switch (e->tag) {
    case T_ARC: return eval2(e->a, rt, v, ctx, cache);                    /* :633 */
    case T_X: return v[0];
    case T_Y: return v[1];
    case T_TAU: return 6.283185307179586;                                 /* :636 */
    case T_E: return 2.718281828459045;                                   /* :637 */
    case T_VAR: return cache_val(cache, rt, v, e->n, ctx).v;              /* :638 */
    case T_NAT: return (double)e->n;                                      /* :639 u64 as f64 */
    case T_NEG: return -eval2(e->a, rt, v, ctx, cache);
    case T_ABS: return fabs(eval2(e->a, rt, v, ctx, cache));
    case T_RECIP: return 1.0 / eval2(e->a, rt, v, ctx, cache);            /* f64::recip */
    case T_SQRT: return sqrt(eval2(e->a, rt, v, ctx, cache));
    case T_STEP: {                                                        /* :644-647 */
        double s = eval2(e->a, rt, v, ctx, cache);
        if (rt->probe) probe_step(e->a, s, rt, v, ctx, cache);
        return (s >= 0.0) ? 1.0 : 0.0;
    }
    case T_SIN: return sin(eval2(e->a, rt, v, ctx, cache));               /* platform libm */
    case T_EXP: return exp(eval2(e->a, rt, v, ctx, cache));
    case T_LN: return log(eval2(e->a, rt, v, ctx, cache));
    case T_ADD: { double a = eval2(e->a, rt, v, ctx, cache); double b = eval2(e->b, rt, v, ctx, cache); return a + b; }
    case T_MUL: { double a = eval2(e->a, rt, v, ctx, cache); double b = eval2(e->b, rt, v, ctx, cache); return a * b; }
    case T_MAX: { double a = eval2(e->a, rt, v, ctx, cache); double b = eval2(e->b, rt, v, ctx, cache); return rust_max(a, b); }
    case T_MIN: { double a = eval2(e->a, rt, v, ctx, cache); double b = eval2(e->b, rt, v, ctx, cache); return rust_min(a, b); }
    case T_LET: return eval2(e->a, rt, v, e->ctx, cache);                 /* :659-662 ctx replaced */
    case T_DECOR: return eval2(e->a, rt, v, ctx, cache);                  /* :663 */
    case T_APP: {                                                         /* :664-668 */
        if (e->app_id >= rt->n_functions) {                               /* Rust: index panic */
            rt->fault = 1;
            (void)eval2(e->a, rt, v, ctx, cache); (void)eval2(e->b, rt, v, ctx, cache);
            return NAN;
        }
        double a = eval2(e->a, rt, v, ctx, cache);
        double b = eval2(e->b, rt, v, ctx, cache);
        return texture_fn(rt->tex, e->app_id, a, b);
    }
    }
    return NAN;
}

/* Diagnostic only (tests attribute a 0<->255 difference between two renderers to a `step` whose
 * argument is a rounding error away from zero): how far is this step's argument from
 * zero, in units in the last place of the larger of the two terms whose sum it is?  Looks through
 * Arc/Decor/Neg and through a Var to its definition; an argument that is not a sum is measured
 * against itself (margin 2^52: never "close").  Not part of the restated algorithm. */
static void probe_step(const Expr *arg, double s, Rt *rt, const double v[2], const Ctx *ctx, Cache *cache) {
    const Expr *t = arg;
    for (int hops = 0; hops < 64; hops++) {
        if (t->tag == T_ARC || t->tag == T_DECOR || t->tag == T_NEG) { t = t->a; continue; }
        if (t->tag == T_VAR) {
            const Expr *def = NULL;
            for (uint64_t i = 0; i < ctx->n; i++) if (ctx->ids[i] == t->n) { def = ctx->defs[i]; break; }
            if (!def) break;
            t = def;
            continue;
        }
        break;
    }
    double scale = fabs(s);
    if (t->tag == T_ADD) {
        Rt q = *rt; q.probe = NULL;
        double a = eval2(t->a, &q, v, ctx, cache), b = eval2(t->b, &q, v, ctx, cache);
        scale = fmax(fabs(a), fabs(b));
    }
    double margin = (s != s) ? INFINITY : ((scale > 0.0) ? fabs(s) / (scale * 0x1p-52) : 0.0);
    if (margin < *rt->probe) *rt->probe = margin;
}

static int dep_x(const Expr *e, Rt *rt, const double v[2], const Ctx *ctx, Cache *cache) {
    switch (e->tag) {
    case T_ARC: return dep_x(e->a, rt, v, ctx, cache);
    case T_X: return 1;
    case T_Y: case T_TAU: case T_E: case T_NAT: return 0;
    case T_VAR: return cache_val(cache, rt, v, e->n, ctx).depx;
    case T_NEG: case T_ABS: case T_RECIP: case T_SQRT:
    case T_STEP: case T_SIN: case T_EXP: case T_LN:
        return dep_x(e->a, rt, v, ctx, cache);
    case T_ADD: case T_MUL: case T_MAX: case T_MIN: case T_APP: {
        int a = dep_x(e->a, rt, v, ctx, cache);
        int b = dep_x(e->b, rt, v, ctx, cache);
        return a || b;
    }
    case T_LET: return dep_x(e->a, rt, v, e->ctx, cache);
    case T_DECOR: return dep_x(e->a, rt, v, ctx, cache);
    }
    return 0;
}

/* `f64 as u8` (src/render.rs:26-28): truncate toward zero, saturate to [0,255], NaN -> 0. */
static inline uint8_t f64_as_u8(double v) {
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

/* ------------------------------------------------------------------------------------------ */
/* Scene handle + exported API (ctypes-friendly).                                              */

typedef struct mo_scene {
    Arena ar;
    uint32_t size[2];
    int legacy;
    Expr *raw[3];          /* as parsed                       */
    Expr *color[3];        /* after var_fixer::fix_color      */
    Textures tex;
} mo_scene;

typedef struct { mo_scene *s; const uint8_t *p; size_t len; int ok; } OpenJob;

/* Layout plausibility: every Var must be bound by an enclosing Let.  A legacy file can by accident
 * also decode under HEAD numbering (each tag means the previous variant: Nat reads as Var, ...);
 * such a mis-decode leaves unbound variables, which no file written by `save` contains. */
typedef struct Scope { const Ctx *ctx; const struct Scope *up; } Scope;
static int vars_bound(const Expr *e, const Scope *sc) {
    if (e->tag == T_VAR) {
        for (const Scope *s = sc; s; s = s->up)
            for (uint64_t i = 0; i < s->ctx->n; i++) if (s->ctx->ids[i] == e->n) return 1;
        return 0;
    }
    if (e->tag == T_LET) {
        Scope in = {e->ctx, sc};
        for (uint64_t i = 0; i < e->ctx->n; i++) if (!vars_bound(e->ctx->defs[i], &in)) return 0;
        return vars_bound(e->a, &in);
    }
    if (e->a && !vars_bound(e->a, sc)) return 0;
    if (e->b && !vars_bound(e->b, sc)) return 0;
    return 1;
}

static int try_parse(mo_scene *s, const uint8_t *p, size_t len, int legacy) {
    Arena ar = {0};
    Rd r; r.p = p; r.end = p + len; r.legacy = legacy; r.err = 0; r.ar = &ar; r.depth = 0;
    uint32_t w = rd_u32(&r), h = rd_u32(&r);
    Expr *c[3];
    for (int i = 0; i < 3; i++) c[i] = rd_expr(&r);
    if (r.err || r.p != r.end) { arena_free(&ar); return 0; }   /* must consume the whole file */
    for (int i = 0; i < 3; i++) if (!vars_bound(c[i], NULL)) { arena_free(&ar); return 0; }
    s->ar = ar; s->size[0] = w; s->size[1] = h; s->legacy = legacy;
    for (int i = 0; i < 3; i++) s->raw[i] = c[i];
    return 1;
}

/* var_fixer::fix_color src/var_fixer.rs:74-82: one VarFixer across R, G, B, empty outer ctx. */
static void fix_color(mo_scene *s) {
    Fixer f; memset(&f, 0, sizeof f); f.ar = &s->ar;
    RenCtx empty = {NULL, 0};
    for (int i = 0; i < 3; i++) s->color[i] = fix(&f, s->raw[i], &empty);
    free(f.ids.keys); free(f.ids.vals);
}

static void *open_thread(void *arg) {
    OpenJob *j = (OpenJob *)arg;
    /* HEAD layout first, then legacy; accept the one that parses to exact EOF with all variables
     * bound. */
    j->ok = try_parse(j->s, j->p, j->len, 0) || try_parse(j->s, j->p, j->len, 1);
    if (j->ok) fix_color(j->s);
    return NULL;
}

/* Deep trees recurse deeply: run tree walks on threads with a large stack. */
static int run_big_stack(void *(*fn)(void *), void *arg, pthread_t *out) {
    pthread_attr_t at; pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, (size_t)1 << 30);
    int rc = pthread_create(out, &at, fn, arg);
    pthread_attr_destroy(&at);
    return rc;
}

/* maray::open src/lib.rs:1227-1235 (bytes instead of a path) + fix_color (src/render.rs:14,51,117). */
mo_scene *mo_open(const uint8_t *bytes, size_t len) {
    mo_scene *s = (mo_scene *)calloc(1, sizeof(mo_scene));
    OpenJob j = {s, bytes, len, 0};
    pthread_t th;
    if (run_big_stack(open_thread, &j, &th)) { free(s); return NULL; }
    pthread_join(th, NULL);
    if (!j.ok) { free(s); return NULL; }
    return s;
}

/* maray::save (src/lib.rs:1216-1224) of the scene AFTER var_fixer::fix_color, HEAD layout: lets the
 * tests compare the fixed trees with the expectations of the reference's own test_var_fixer
 * (src/lib.rs:1508-1691).  Returns the number of bytes the encoding needs; writes at most cap. */
typedef struct { uint8_t *p; size_t cap, n; } Wr;
static void wr_bytes(Wr *w, const void *src, size_t k) {
    if (w->n + k <= w->cap) memcpy(w->p + w->n, src, k);
    w->n += k;
}
static void wr_u32(Wr *w, uint32_t v) { wr_bytes(w, &v, 4); }   /* little-endian host, like the reader */
static void wr_u64(Wr *w, uint64_t v) { wr_bytes(w, &v, 8); }
static void wr_expr(Wr *w, const Expr *e) {
    wr_u32(w, e->tag);
    switch (e->tag) {
    case T_VAR: case T_NAT: wr_u64(w, e->n); break;
    case T_LET:
        wr_u64(w, e->ctx->n);
        for (uint64_t i = 0; i < e->ctx->n; i++) { wr_u64(w, e->ctx->ids[i]); wr_expr(w, e->ctx->defs[i]); }
        wr_expr(w, e->a);
        break;
    case T_DECOR: wr_expr(w, e->a); wr_u64(w, 0); break;          /* tokens were dropped by the reader */
    case T_APP: wr_u32(w, e->app_id); wr_expr(w, e->a); wr_expr(w, e->b); break;
    default:
        if (e->a) wr_expr(w, e->a);
        if (e->b) wr_expr(w, e->b);
    }
}
size_t mo_fixed_bytes(const mo_scene *s, uint8_t *buf, size_t cap) {
    Wr w = {buf, buf ? cap : 0, 0};
    wr_u32(&w, s->size[0]); wr_u32(&w, s->size[1]);
    for (int i = 0; i < 3; i++) wr_expr(&w, s->color[i]);
    return w.n;
}

void mo_close(mo_scene *s) {
    if (!s) return;
    for (uint32_t i = 0; i < s->tex.n; i++) free(s->tex.data[i]);
    free(s->tex.data); free(s->tex.w); free(s->tex.h);
    arena_free(&s->ar);
    free(s);
}

void mo_size(const mo_scene *s, uint32_t *w, uint32_t *h) { *w = s->size[0]; *h = s->size[1]; }
int mo_is_legacy_layout(const mo_scene *s) { return s->legacy; }

static uint64_t count_nodes(const Expr *e) {
    uint64_t n = 1;
    if (e->tag == T_LET) for (uint64_t i = 0; i < e->ctx->n; i++) n += count_nodes(e->ctx->defs[i]);
    if (e->a) n += count_nodes(e->a);
    if (e->b) n += count_nodes(e->b);
    return n;
}
typedef struct { const mo_scene *s; int ch; uint64_t out; } CountJob;
static void *count_thread(void *arg) { CountJob *j = (CountJob *)arg; j->out = count_nodes(j->s->raw[j->ch]); return NULL; }
/* Tree node count of one channel as stored on the wire (diagnostic; chess: 29 314). */
uint64_t mo_tree_nodes(const mo_scene *s, int channel) {
    CountJob j = {s, channel, 0}; pthread_t th;
    if (run_big_stack(count_thread, &j, &th)) return 0;
    pthread_join(th, NULL);
    return j.out;
}

/* Textures{images} + textures::functions(n) (src/textures.rs:9-12,54-65; examples/maray.rs:58-69).
 * Copies the RGB8 data. */
int mo_set_textures(mo_scene *s, uint32_t n, const uint8_t *const *rgb, const uint32_t *w, const uint32_t *h) {
    for (uint32_t i = 0; i < s->tex.n; i++) free(s->tex.data[i]);
    free(s->tex.data); free(s->tex.w); free(s->tex.h);
    s->tex.n = n;
    s->tex.data = (uint8_t **)calloc(n ? n : 1, sizeof(uint8_t *));
    s->tex.w = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    s->tex.h = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    for (uint32_t i = 0; i < n; i++) {
        size_t sz = (size_t)w[i] * h[i] * 3;
        s->tex.data[i] = (uint8_t *)malloc(sz ? sz : 1);
        memcpy(s->tex.data[i], rgb[i], sz);
        s->tex.w[i] = w[i]; s->tex.h[i] = h[i];
    }
    return 0;
}

/* ---- Expr::eval analogue (src/lib.rs:617-620): one channel at one point, fresh Cache. ------ */
typedef struct { const mo_scene *s; int ch; double x, y; double out; int fault; } EvalJob;
static void *eval_thread(void *arg) {
    EvalJob *j = (EvalJob *)arg;
    Rt rt = {&j->s->tex, j->s->tex.n * ALIGN, 0, NULL};
    Cache c; cache_init(&c);
    double v[2] = {j->x, j->y};
    j->out = eval2(j->s->color[j->ch], &rt, v, &EMPTY_CTX, &c);
    j->fault = rt.fault;
    cache_free(&c);
    return NULL;
}
double mo_eval(const mo_scene *s, int channel, double x, double y) {
    EvalJob j = {s, channel, x, y, 0.0, 0}; pthread_t th;
    if (run_big_stack(eval_thread, &j, &th)) return NAN;
    pthread_join(th, NULL);
    return j.out;
}

/* Diagnostic: the smallest step margin (see probe_step) over all `step`s the three channels evaluate at
 * pixel (x, y); +inf when no step is evaluated. */
typedef struct { const mo_scene *s; double x, y; double out; } MarginJob;
static void *margin_thread(void *arg) {
    MarginJob *j = (MarginJob *)arg;
    double m = INFINITY;
    Rt rt = {&j->s->tex, j->s->tex.n * ALIGN, 0, &m};
    Cache c; cache_init(&c);
    double v[2] = {j->x, j->y};
    for (int ch = 0; ch < 3; ch++) (void)eval2(j->s->color[ch], &rt, v, &EMPTY_CTX, &c);
    cache_free(&c);
    j->out = m;
    return NULL;
}
double mo_step_margin(const mo_scene *s, double x, double y) {
    MarginJob j = {s, x, y, 0.0}; pthread_t th;
    if (run_big_stack(margin_thread, &j, &th)) return NAN;
    pthread_join(th, NULL);
    return j.out;
}

/* ---- par_gen_to_image restatement (src/render.rs:35-99). ----------------------------------
 * Rows are independent work items (rayon `into_par_iter` over 0..h, :85); each row gets a fresh
 * Cache (:86), every pixel starts with cache.clear_dep_x() (:89), p = [x as f64, y as f64] (:90),
 * channels are evaluated R,G,B through the same cache (:91-93) and cast with `as u8`.
 * The collector thread (:57-83) only moves finished rows into the image; here rows are written
 * in place.  Report::None (no progress callback).  Threads pull rows from a shared counter,
 * which is also the scheduling of wasm_par_gen_to_image (:150-173).
 *
 * A window [x0,x1) x [y0,y1) can be rendered so tests can sample big scenes; every row still
 * starts with an empty cache, and because `Let` values are pure functions of (x,y) the window
 * equals the same region of the full image.  f64_planes (optional) receives the raw channel
 * values, 3 planes of (y1-y0)*(x1-x0) doubles.                                                */
typedef struct {
    const mo_scene *s;
    uint32_t x0, x1;
    const uint32_t *rows; uint32_t n_rows;      /* image rows to render; output row i = rows[i] */
    uint8_t *rgb; double *planes;
    pthread_mutex_t mu; uint32_t next;          /* shared row counter (reference src/render.rs:150) */
    int fault;
} RenderJob;

static void *render_thread(void *arg) {
    RenderJob *j = (RenderJob *)arg;
    const mo_scene *s = j->s;
    Rt rt = {&s->tex, s->tex.n * ALIGN, 0, NULL};
    Cache cache; cache_init(&cache);
    uint32_t ww = j->x1 - j->x0, hh = j->n_rows;
    size_t plane = (size_t)ww * hh;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        uint32_t i = j->next;
        if (i < j->n_rows) j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n_rows) break;
        uint32_t y = j->rows[i];
        cache_clear(&cache);                                         /* fresh Cache per row */
        for (uint32_t x = j->x0; x < j->x1; x++) {
            cache_clear_dep_x(&cache);
            double p[2] = {(double)x, (double)y};
            double r = eval2(s->color[0], &rt, p, &EMPTY_CTX, &cache);
            double g = eval2(s->color[1], &rt, p, &EMPTY_CTX, &cache);
            double b = eval2(s->color[2], &rt, p, &EMPTY_CTX, &cache);
            size_t o = (size_t)i * ww + (x - j->x0);
            if (j->rgb) {
                j->rgb[o * 3 + 0] = f64_as_u8(r);
                j->rgb[o * 3 + 1] = f64_as_u8(g);
                j->rgb[o * 3 + 2] = f64_as_u8(b);
            }
            if (j->planes) {
                j->planes[o] = r; j->planes[plane + o] = g; j->planes[2 * plane + o] = b;
            }
        }
    }
    if (rt.fault) j->fault = 1;
    cache_free(&cache);
    return NULL;
}

/* Renders the listed rows (any order, any subset) of columns [x0,x1).
 * Returns 0, or -1 when the reference would have panicked (App id outside Runtime::functions). */
int mo_render_rows(const mo_scene *s, uint32_t x0, uint32_t x1, const uint32_t *rows, uint32_t n_rows,
                   int nthreads, uint8_t *rgb, double *f64_planes) {
    if (x1 < x0) return -2;
    if (n_rows == 0 || x1 == x0) return 0;
    RenderJob j; memset(&j, 0, sizeof j);
    j.s = s; j.x0 = x0; j.x1 = x1; j.rows = rows; j.n_rows = n_rows; j.rgb = rgb; j.planes = f64_planes;
    pthread_mutex_init(&j.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n_rows) nthreads = (int)n_rows;
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    int started = 0;
    for (int i = 0; i < nthreads; i++) if (!run_big_stack(render_thread, &j, &th[started])) started++;
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&j.mu);
    if (!started) return -3;
    return j.fault ? -1 : 0;
}

int mo_render_window(const mo_scene *s, uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
                     int nthreads, uint8_t *rgb, double *f64_planes) {
    if (x1 < x0 || y1 < y0) return -2;
    uint32_t n = y1 - y0;
    uint32_t *rows = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    for (uint32_t i = 0; i < n; i++) rows[i] = y0 + i;
    int rc = mo_render_rows(s, x0, x1, rows, n, nthreads, rgb, f64_planes);
    free(rows);
    return rc;
}

/* Whole image of size w x h (gen_to_image with an RgbImage of that size, src/lib.rs:1177-1195). */
int mo_render(const mo_scene *s, uint32_t w, uint32_t h, int nthreads, uint8_t *rgb) {
    return mo_render_window(s, 0, w, 0, h, nthreads, rgb, NULL);
}
