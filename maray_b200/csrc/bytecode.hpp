// Bytecode of the interpreter back end: a flat, single-pass register program with one accumulator
// per pixel and numbered value slots (version 2).
//
// Why this shape: the per-pixel slots live in shared memory, and shared-memory capacity and latency
// -- not the FP64 pipe -- bound the interpreter.  Every instruction is
//
//      acc = op(first, second);   if (ST) slot[dst] = acc
//
// where `first` is the accumulator (ACC_A) or slot[a] and `second` is slot[b], constant[b] (B_CONST)
// or the accumulator (FWD_B: b is the slot the previous instruction just stored).  Keeping chains in
// the accumulator saves slot traffic; naming both operands lets the kernel fetch the NEXT
// instruction's operands while the current one executes (the loop is latency-bound otherwise).
//
// Instruction word (64 bit):
//   bits  0.. 7  opcode (BcOp)
//   bits  8..15  flags (BC_F_*)
//   bits 16..31  dst slot (BC_TEX: texture*4 + channel instead; TEX never stores)
//   bits 32..47  a: slot index of the first operand (ignored with ACC_A)
//   bits 48..63  b: slot or constant index of the second operand
// SWAP evaluates op(second, first) -- f64::max/min and App are not symmetric (NaN / +-0 / x,y).
// Slot 0 holds X and slot 1 holds Y (`x as f64`, `y as f64`) when the program starts.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "program.hpp"

namespace maray {

enum BcOp : uint8_t {
    BC_END = 0,
    BC_MOV,                                            // acc = first (or second with SWAP)
    BC_ADD, BC_MUL, BC_MAX, BC_MIN,                    // acc = first op second
    BC_NEG, BC_ABS, BC_RECIP, BC_SQRT, BC_STEP, BC_SIN, BC_EXP, BC_LN,   // acc = f(first)
    BC_TEX,                                            // acc = texture(dst field)(x = first, y = second)
    BC_OUT_R, BC_OUT_G, BC_OUT_B,                      // channel value = first
    BC_COUNT
};

constexpr uint32_t BC_F_STORE = 1u;      // slot[dst] = acc after the operation
constexpr uint32_t BC_F_ACC_A = 2u;      // first operand is the accumulator
constexpr uint32_t BC_F_SWAP = 4u;       // compute op(second, first)
constexpr uint32_t BC_F_B_CONST = 8u;    // second operand is constant[b]
constexpr uint32_t BC_F_FWD_B = 16u;     // second operand is the accumulator (slot b was stored by the previous instruction)
// Row-uniform slots (optional form, compile_bytecode(.., row_uniform = true)): a value that depends on y
// only is the same for every pixel of a row, so when every block of a launch lies inside one row it
// needs ONE word per block, not one per thread.  Uniform slots are numbered separately and never
// recycled: every warp computes and stores every such value itself (same bits), so a warp only ever
// reads what it has written and no barrier is needed.
constexpr uint32_t BC_F_A_UNI = 32u;     // first operand is uniform slot a
constexpr uint32_t BC_F_B_UNI = 64u;     // second operand is uniform slot b
constexpr uint32_t BC_F_ST_UNI = 128u;   // the store (BC_F_STORE) goes to uniform slot dst

struct Bytecode {
    std::vector<uint64_t> code;      // ends with BC_END
    std::vector<double> consts;
    uint32_t n_slots = 2;            // per-pixel slots, including X and Y
    uint32_t n_uniform = 0;          // per-block slots (row-uniform form only)
};

inline uint64_t bc_encode(BcOp op, uint32_t flags, uint32_t dst, uint32_t a, uint32_t b) {
    return uint64_t(op) | (uint64_t(flags & 0xff) << 8) | (uint64_t(dst & 0xffff) << 16) | (uint64_t(a & 0xffff) << 32) |
           (uint64_t(b & 0xffff) << 48);
}

bool compile_bytecode(const Program& prog, Bytecode* out, std::string* err, bool row_uniform = false);

}  // namespace maray
