// Bytecode of the interpreter back end (version 3): a flat, single-pass program with TWO accumulators
// and two slot files, laid out for a launch in which every block lies inside one image row.
//
// Why this shape.  A block of B threads renders B*P consecutive pixels of ONE row, so a value that does
// not depend on x ("row-uniform": everything made of y and constants -- the reference keeps exactly these
// across a row in its Cache, reference src/cache.rs:18-20 + Expr::dep_x src/lib.rs:675-706) is one number
// per block.  Such values are computed once per thread as a SCALAR (one FP64 instruction per warp instead
// of P) and live in a small per-block scalar file next to the constants; values that depend on x are
// WIDE (P per thread) and live in the per-thread slot file, which is what bounds occupancy.  On the
// shipped chess scene this takes the per-pixel slot file from 86 to ~30 values.
//
// Every instruction is
//
//      acc  = op(first, second);   [wide[dst] = acc]       (wide shape:   P values per thread)
//      sacc = op(first, second);   [scal[dst] = sacc]      (scalar shape: 1 value per block)
//
// and each operand is one of four kinds:
//      A  the wide accumulator            W  wide slot (per pixel)
//      T  the scalar accumulator          S  scalar file entry (constant pool, then row-uniform slots)
// A scalar instruction reads T/S only.  There are no operand-order tricks: f64::max/min and App are not
// symmetric (NaN, +-0, x/y), so both operands can be of any kind and keep the scene's order.
//
// Instruction word (64 bit):
//   bits  0.. 7  handler id (BcHandler): shape, operation and -- for the hot wide operations -- both
//                operand kinds, so the kernel's dispatch is ONE indexed jump into a specialised body
//   bits  8..15  flags: BC_F_STORE, and the operand kinds again (read by the generic bodies)
//   bits 16..31  dst slot (wide or scalar file by shape).  TEX: texture*4 + channel (TEX never stores)
//   bits 32..47  a: wide slot / scalar index of the first operand (ignored for A/T)
//   bits 48..63  b: same for the second operand
// Wide slot 0 holds X (`x as f64`) and scalar entry n_consts holds Y when the program starts
// (row-uniform form); in the all-wide form (row_uniform = false) wide slot 1 holds Y.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "program.hpp"

namespace maray {

enum BcKind : uint8_t { BC_K_A = 0, BC_K_W = 1, BC_K_S = 2, BC_K_T = 3 };

enum BcOp : uint8_t {
    BC_END = 0,
    BC_MOV,                                            // acc = first
    BC_ADD, BC_MUL, BC_MAX, BC_MIN,                    // acc = first op second
    BC_NEG, BC_ABS, BC_RECIP, BC_SQRT, BC_STEP, BC_SIN, BC_EXP, BC_LN,   // acc = f(first)
    BC_TEX,                                            // acc = texture(dst field)(x = first, y = second)
    BC_OUT_R, BC_OUT_G, BC_OUT_B,                      // channel value = first (accumulators unchanged)
    BC_COUNT
};

// Handler ids.  Wide operations with both kinds in the id (specialised bodies):
//   binary  ADD/MUL/MAX/MIN : BC_H_BIN + (op - BC_ADD) * 16 + ka * 4 + kb
//   unary   NEG..LN, MOV    : BC_H_UN  + u * 4 + ka          (u = op - BC_NEG; MOV is u = 8)
//   OUT_R/G/B               : BC_H_OUT + c * 4 + ka
//   TEX (wide)              : BC_H_TEX   (kinds from the flags)
// Scalar shape (operands are S or T only; t = 1 for T):
//   binary                  : BC_H_SBIN + (op - BC_ADD) * 4 + ta * 2 + tb
//   unary, MOV              : BC_H_SUN  + u * 2 + ta
//   TEX                     : BC_H_STEX  (kinds from the flags)
// The ids are dense on purpose: the kernel's dispatch is a jump table over 0 .. BC_H_COUNT-1
// (tools/gen_interp_dispatch.py holds the same numbers).
constexpr uint32_t BC_H_END = 0;
constexpr uint32_t BC_H_YIELD = 1;                     // device stream only: last word of every staged chunk but the final one
constexpr uint32_t BC_H_BIN = 16;                      // 16 .. 79
constexpr uint32_t BC_H_UN = 80;                       // 80 .. 115
constexpr uint32_t BC_H_OUT = 116;                     // 116 .. 127
constexpr uint32_t BC_H_TEX = 128;
constexpr uint32_t BC_H_SCALAR = 144;                  // first scalar-shape id
constexpr uint32_t BC_H_SBIN = 144;                    // 144 .. 159
constexpr uint32_t BC_H_SUN = 160;                     // 160 .. 177
constexpr uint32_t BC_H_STEX = 178;
// Wide binary operation whose accumulator operand is NEGATED first (fuses `neg` into its only consumer; the
// arithmetic is unchanged: x + (-acc), (-acc) * y, max/min of the negated value -- what the scene computes):
//   BC_H_BINN + (op - BC_ADD) * 6 + c,  c = 0..2: (A, W|S|T), 3..5: (W|S|T, A);  flags carry BC_F_NEG_ACC.
constexpr uint32_t BC_H_BINN = 179;                    // 179 .. 202
constexpr uint32_t BC_H_COUNT = 203;

constexpr uint32_t BC_F_STORE = 1u;                    // store the result to slot dst of the shape's file
constexpr uint32_t BC_F_NEG_ACC = 2u;                  // the accumulator operand is negated (BC_H_BINN handlers)
constexpr uint32_t BC_F_KA_SHIFT = 2, BC_F_KB_SHIFT = 4;   // operand kinds (BcKind), two bits each

struct Bytecode {
    std::vector<uint64_t> code;      // ends with BC_H_END
    std::vector<double> consts;      // scalar file entries 0 .. consts.size()-1
    uint32_t n_wide = 1;             // wide slots, including X (and Y in the all-wide form)
    uint32_t n_uniform = 0;          // row-uniform scalar slots, including Y (0 in the all-wide form)
    bool row_uniform = true;
};

inline uint64_t bc_encode(uint32_t handler, uint32_t flags, uint32_t dst, uint32_t a, uint32_t b) {
    return uint64_t(handler & 0xff) | (uint64_t(flags & 0xff) << 8) | (uint64_t(dst & 0xffff) << 16) |
           (uint64_t(a & 0xffff) << 32) | (uint64_t(b & 0xffff) << 48);
}
inline uint32_t bc_handler(BcOp op, bool scalar_shape, uint32_t ka, uint32_t kb, bool neg_acc = false) {
    if (op == BC_END) return BC_H_END;
    if (neg_acc) return BC_H_BINN + (op - BC_ADD) * 6 + (ka == BC_K_A ? kb - 1 : 3 + (ka - 1));
    if (scalar_shape) {
        const uint32_t ta = ka == BC_K_T ? 1u : 0u, tb = kb == BC_K_T ? 1u : 0u;
        if (op >= BC_ADD && op <= BC_MIN) return BC_H_SBIN + (op - BC_ADD) * 4 + ta * 2 + tb;
        if (op >= BC_NEG && op <= BC_LN) return BC_H_SUN + (op - BC_NEG) * 2 + ta;
        if (op == BC_MOV) return BC_H_SUN + 8 * 2 + ta;
        return BC_H_STEX;
    }
    if (op >= BC_ADD && op <= BC_MIN) return BC_H_BIN + (op - BC_ADD) * 16 + ka * 4 + kb;
    if (op >= BC_NEG && op <= BC_LN) return BC_H_UN + (op - BC_NEG) * 4 + ka;
    if (op == BC_MOV) return BC_H_UN + 8 * 4 + ka;
    if (op >= BC_OUT_R && op <= BC_OUT_B) return BC_H_OUT + (op - BC_OUT_R) * 4 + ka;
    return BC_H_TEX;
}

// row_uniform = false compiles the all-wide form (every value per pixel; used when a program has more
// row-uniform values than the scalar file holds).
bool compile_bytecode(const Program& prog, Bytecode* out, std::string* err, bool row_uniform = true);

// The program as the kernel reads it for one launch shape:
//  * every wide slot index (operands of kind W and the dst of a storing wide instruction) multiplied by
//    slot16 = P * B / 2, the size of one wide slot in 16-byte units, so an operand address is one shift-add;
//  * cut into chunks of exactly kBcChunk words, the unit the kernel stages into shared memory: every chunk
//    but the last ends with a BC_H_YIELD word, the last is padded with BC_H_END -- the kernel's inner loop
//    needs no bounds check, it leaves at the YIELD / END word.
// Empty + err when a field would overflow.
constexpr uint32_t kBcChunk = 256;
std::vector<uint64_t> bytecode_for_launch(const Bytecode& bc, uint32_t slot16, std::string* err);

}  // namespace maray
