// `.maray` loader: see expr.hpp.  Wire layout per SURVEY.md Appendix A (bincode 1.3.3 defaults:
// little-endian, fixed-width integers, u32 enum tags, u64 lengths, Box/Arc transparent).
#include "expr.hpp"

#include <pthread.h>

#include <cstdlib>
#include <cstring>
#include <new>

namespace maray {

Arena::~Arena() {
    Block* b = head_;
    while (b) { Block* n = b->next; std::free(b); b = n; }
}

void* Arena::alloc(size_t bytes) {
    bytes = (bytes + 15) & ~size_t(15);
    if (!head_ || head_->used + bytes > head_->cap) {
        size_t cap = bytes > (size_t(1) << 20) ? bytes : (size_t(1) << 20);
        Block* b = static_cast<Block*>(std::malloc(sizeof(Block) + cap));
        if (!b) throw std::bad_alloc();
        b->next = head_; b->used = 0; b->cap = cap;
        head_ = b;
    }
    char* p = reinterpret_cast<char*>(head_ + 1) + head_->used;
    head_->used += bytes;
    std::memset(p, 0, bytes);
    return p;
}

namespace {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool legacy;
    Arena* arena;
    const char* err = nullptr;
    uint32_t depth = 0;
    uint64_t nodes = 0;

    bool need(size_t n) {
        if (err) return false;
        if (size_t(end - p) < n) { err = "unexpected end of file"; return false; }
        return true;
    }
    uint32_t u32() { if (!need(4)) return 0; uint32_t v; std::memcpy(&v, p, 4); p += 4; return v; }
    uint64_t u64() { if (!need(8)) return 0; uint64_t v; std::memcpy(&v, p, 8); p += 8; return v; }

    const Expr* expr();
    void token();
};

// Token (reference src/token.rs:11-38).  Decor is transparent for evaluation (reference
// src/lib.rs:663), so tokens are only skipped.
void Reader::token() {
    uint32_t t = u32();
    if (err) return;
    if (t == 0) { (void)expr(); }
    else if (t == 1) {
        uint64_t len = u64();
        if (err) return;
        if (uint64_t(end - p) < len) { err = "string token runs past end of file"; return; }
        p += len;
    } else if (t > 12) err = "invalid Token variant";
}

const Expr* Reader::expr() {
    if (err) return nullptr;
    if (++depth > 4000000) { err = "expression nesting too deep"; return nullptr; }
    uint32_t t = u32();
    if (err) return nullptr;
    if (legacy) t += 1;   // the legacy numbering has no Arc variant
    if (t >= T_COUNT || (legacy && t == T_ARC)) { err = "invalid Expr variant"; return nullptr; }
    Expr* e = arena->make<Expr>();
    e->tag = Tag(t);
    nodes++;
    switch (t) {
    case T_ARC: e->a = expr(); break;
    case T_X: case T_Y: case T_TAU: case T_E: break;
    case T_VAR: case T_NAT: e->n = u64(); break;
    case T_NEG: case T_ABS: case T_RECIP: case T_SQRT:
    case T_STEP: case T_SIN: case T_EXP: case T_LN:
        e->a = expr(); break;
    case T_ADD: case T_MUL: case T_MAX: case T_MIN:
        e->a = expr(); e->b = expr(); break;
    case T_LET: {
        uint64_t n = u64();
        if (err) return nullptr;
        if (n > uint64_t(end - p) / 12) { err = "Let length exceeds file size"; return nullptr; }
        LetVar* vars = arena->make<LetVar>(n ? n : 1);
        for (uint64_t i = 0; i < n && !err; i++) {
            vars[i].id = u64();
            vars[i].def = expr();
        }
        e->vars = vars; e->n_vars = n;
        e->a = expr();
        break;
    }
    case T_DECOR: {
        e->a = expr();
        uint64_t n = u64();
        for (uint64_t i = 0; i < n && !err; i++) token();
        break;
    }
    case T_APP:
        e->app_id = u32();
        e->a = expr(); e->b = expr();
        break;
    }
    depth--;
    return err ? nullptr : e;
}

struct Scope { const Expr* let; const Scope* up; };

bool vars_bound(const Expr* e, const Scope* sc) {
    if (e->tag == T_VAR) {
        for (const Scope* s = sc; s; s = s->up)
            for (uint64_t i = 0; i < s->let->n_vars; i++)
                if (s->let->vars[i].id == e->n) return true;
        return false;
    }
    if (e->tag == T_LET) {
        Scope in{e, sc};
        for (uint64_t i = 0; i < e->n_vars; i++) if (!vars_bound(e->vars[i].def, &in)) return false;
        return vars_bound(e->a, &in);
    }
    if (e->a && !vars_bound(e->a, sc)) return false;
    if (e->b && !vars_bound(e->b, sc)) return false;
    return true;
}

const char* try_layout(const uint8_t* bytes, size_t len, bool legacy, Scene* out) {
    std::unique_ptr<Arena> arena(new Arena());
    Reader r{bytes, bytes + len, legacy, arena.get()};
    uint32_t w = r.u32(), h = r.u32();
    const Expr* c[3] = {nullptr, nullptr, nullptr};
    uint64_t counts[3] = {0, 0, 0};
    for (int i = 0; i < 3 && !r.err; i++) {
        uint64_t before = r.nodes;
        c[i] = r.expr();
        counts[i] = r.nodes - before;
    }
    if (r.err) return r.err;
    if (r.p != r.end) return "trailing bytes after the third channel";
    // A legacy file can decode by accident under HEAD numbering (every tag then means the previous
    // variant: Nat reads as Var ...).  Such a mis-decode leaves unbound variables.
    for (int i = 0; i < 3; i++) if (!vars_bound(c[i], nullptr)) return "unbound variable";
    out->size[0] = w; out->size[1] = h;
    for (int i = 0; i < 3; i++) { out->color[i] = c[i]; out->tree_nodes[i] = counts[i]; }
    out->legacy_layout = legacy;
    out->arena = std::move(arena);
    return nullptr;
}

struct ParseJob { const uint8_t* bytes; size_t len; Scene* out; std::string* err; bool ok; };

void parse_job(void* arg) {
    ParseJob* j = static_cast<ParseJob*>(arg);
    try {
        const char* e_head = try_layout(j->bytes, j->len, false, j->out);
        if (!e_head) { j->ok = true; return; }
        const char* e_legacy = try_layout(j->bytes, j->len, true, j->out);
        if (!e_legacy) { j->ok = true; return; }
        if (j->err) *j->err = std::string("not a .maray file (current layout: ") + e_head + "; legacy layout: " + e_legacy + ")";
    } catch (const std::bad_alloc&) {
        if (j->err) *j->err = "out of memory while parsing";
    }
    j->ok = false;
}

struct Thunk { void (*fn)(void*); void* arg; };
void* thunk_main(void* p) { Thunk* t = static_cast<Thunk*>(p); t->fn(t->arg); return nullptr; }

}  // namespace

bool run_with_big_stack(void (*fn)(void*), void* arg) {
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, size_t(1) << 30);
    Thunk t{fn, arg};
    pthread_t th;
    // No fallback to the caller's stack: the loader accepts trees nested millions deep, which the
    // default 8 MiB cannot survive -- failing is better than a stack overflow.
    const bool started = pthread_create(&th, &at, thunk_main, &t) == 0;
    if (started) pthread_join(th, nullptr);
    pthread_attr_destroy(&at);
    return started;
}

bool parse_maray(const uint8_t* bytes, size_t len, Scene* out, std::string* err) {
    if (len < 8) { if (err) *err = "not a .maray file (shorter than the size header)"; return false; }
    ParseJob j{bytes, len, out, err, false};
    if (!run_with_big_stack(parse_job, &j)) {
        if (err) *err = "cannot start the parser thread (1 GiB stack unavailable)";
        return false;
    }
    return j.ok;
}

}  // namespace maray
