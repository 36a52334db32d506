// Bytecode of the interpreter back end: a flat, static-single-pass register program with one
// accumulator per pixel and numbered value slots.
//
// Why an accumulator machine: the per-pixel slots live in shared memory, and shared-memory
// bandwidth (128 B/clk/SM), not the FP64 pipe, bounds a three-address design (24 B of slot traffic
// per FP64 operation).  Keeping the running value in a register and letting every instruction
// optionally store its result cuts that to 8-16 B (DESIGN.md "Interpreter kernel").
//
// Instruction word (64 bit):
//   bits  0.. 7  opcode (BcOp)
//   bit   8      store flag: after the operation, slot[dst] = acc
//   bits 16..31  dst slot
//   bits 32..63  operand: slot index (*_S), constant-pool index (*_K),
//                or for BC_TEX*: low 16 bits slot index, high 16 bits texture*4 + channel
// Slot 0 holds X and slot 1 holds Y (`x as f64`, `y as f64`) when the program starts.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "program.hpp"

namespace maray {

enum BcOp : uint8_t {
    BC_END = 0,
    BC_LD_S, BC_LD_K,                                   // acc = operand
    BC_NEG, BC_ABS, BC_RECIP, BC_SQRT, BC_STEP, BC_SIN, BC_EXP, BC_LN,   // acc = f(acc)
    BC_ADD_S, BC_ADD_K, BC_MUL_S, BC_MUL_K,             // acc = acc op operand
    BC_MAX_S, BC_MAX_K, BC_MAXR_S, BC_MAXR_K,           // MAX: max(acc, operand); MAXR: max(operand, acc)
    BC_MIN_S, BC_MIN_K, BC_MINR_S, BC_MINR_K,
    BC_TEX_S, BC_TEXR_S,                                // TEX: tex(x=operand, y=acc); TEXR: tex(x=acc, y=operand)
    BC_OUT_R, BC_OUT_G, BC_OUT_B,                       // channel value = acc
    BC_COUNT
};

constexpr uint32_t BC_FLAG_STORE = 1u << 8;

struct Bytecode {
    std::vector<uint64_t> code;      // ends with BC_END
    std::vector<double> consts;
    uint32_t n_slots = 2;            // including X and Y
};

inline uint64_t bc_encode(BcOp op, uint32_t operand) { return uint64_t(op) | (uint64_t(operand) << 32); }

bool compile_bytecode(const Program& prog, Bytecode* out, std::string* err);

}  // namespace maray
