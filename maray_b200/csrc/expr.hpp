// Expr tree + `.maray` loader (host side of the CUDA render path).
//
// Mirrors the reference's IR and `open`:
//   Expr enum          reference src/lib.rs:101-149
//   Context            reference src/lib.rs:51-55
//   open() / bincode   reference src/lib.rs:1227-1235 (wire layout: SURVEY.md Appendix A)
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace maray {

enum Tag : uint32_t {   // HEAD variant order
    T_ARC = 0, T_X, T_Y, T_TAU, T_E, T_VAR, T_NAT,
    T_NEG, T_ABS, T_RECIP, T_SQRT, T_STEP, T_SIN, T_EXP, T_LN,
    T_ADD, T_MUL, T_MAX, T_MIN, T_LET, T_DECOR, T_APP, T_COUNT
};

struct Expr;
struct LetVar { uint64_t id; const Expr* def; };

struct Expr {
    Tag tag;
    uint32_t app_id = 0;        // App
    uint64_t n = 0;             // Var id / Nat value
    const Expr* a = nullptr;    // unary operand, first binary operand, Let body, Decor/Arc inner
    const Expr* b = nullptr;    // second binary operand
    const LetVar* vars = nullptr;   // Let context
    uint64_t n_vars = 0;
};

// Bump allocator owning every node of a scene.
class Arena {
public:
    Arena() = default;
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;
    ~Arena();
    void* alloc(size_t bytes);
    template <class T> T* make(size_t count = 1) { return static_cast<T*>(alloc(sizeof(T) * count)); }
private:
    struct Block { Block* next; size_t used, cap; };
    Block* head_ = nullptr;
};

struct Scene {
    uint32_t size[2] = {0, 0};
    const Expr* color[3] = {nullptr, nullptr, nullptr};   // R, G, B
    bool legacy_layout = false;
    uint64_t tree_nodes[3] = {0, 0, 0};
    std::unique_ptr<Arena> arena;
};

// Parses bincode(([u32;2],[Expr;3])).  Accepts the HEAD layout and the legacy (pre-`Arc`) layout
// used by data/chess.maray; the accepted layout is the one that consumes the whole buffer with
// every variable bound.  Returns false and fills `err` on malformed input.
bool parse_maray(const uint8_t* bytes, size_t len, Scene* out, std::string* err);

// Runs fn(arg) on a thread with a 1 GiB (virtual) stack: scene trees can be very deep and the
// loader/lowering are recursive.  Returns false (fn not run) when the thread cannot be created; fn must
// not let an exception escape.
bool run_with_big_stack(void (*fn)(void*), void* arg);

}  // namespace maray
