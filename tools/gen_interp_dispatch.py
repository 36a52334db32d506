"""Generates maray_b200/csrc/interp_dispatch.inc: the interpreter kernel's dispatch and its hot handler
bodies as inline PTX, one macro per pixels-per-thread count P in {1, 2, 4}.

Why PTX: nvcc (12.9) lowers a dense C++ `switch` to a tree of compare-and-branch pairs -- 13 branch
instructions per bytecode instruction, measured with ncu on the first build of this kernel
(profiles/r02_chess1k_interp_v3tree_ncu_full_summary.txt) -- and never to an indexed branch.  PTX has one
(`brx.idx` over a `.branchtargets` table; SASS: LDC from the table + BRX), but its targets must be labels of
the same asm block, so the bodies live here too: each is specialised on the operation and on where both
operands come from (bytecode.hpp: A wide accumulator, W wide slot, S scalar file, T scalar accumulator), i.e.
it is its loads plus P FP64 instructions.  The asm block is the interpreter's inner loop: every body first turns the fields of its instruction word into addresses, then fetches the next word
over the current one (that latency hides behind the body's own loads and arithmetic), does its work, stores
if it is a storing body (the store flag is part of the table index, so there is no flag test), and jumps back
to the one shared dispatch site -- ~13 SASS instructions per bytecode instruction, two of them branches.
Control returns to C++ only at an instruction without a body here (recip, sqrt, sin, exp, ln, texture fetch,
channel outputs: the kernel's C++ switch runs that one and re-enters) and at the YIELD / END words that close
every staged chunk, so there is no bounds check either.

Operation semantics are those of device_sem.cuh (reference src/lib.rs:632-669):
  add/mul    add.rn.f64 / mul.rn.f64 -- never fused
  max(x, y)  take y when (y > x) || isnan(x), else x   (f64::max as compiled for x86-64; first operand wins ties)
  min(x, y)  take y when (y < x) || isnan(x), else x
  step(x)    x >= 0.0 ? 1.0 : 0.0   (ordered compare: NaN -> 0, -0.0 -> 1)
  neg / abs  sign-bit operations

usage: python tools/gen_interp_dispatch.py > maray_b200/csrc/interp_dispatch.inc
"""
import sys

# handler ids: keep in step with csrc/bytecode.hpp
H_BIN, H_UN, H_OUT, H_TEX = 16, 80, 116, 128
H_SBIN, H_SUN, H_STEX, H_BINN, H_COUNT = 144, 160, 178, 179, 203
A, W, S, T = 0, 1, 2, 3
BIN_OPS = ["add", "mul", "max", "min"]
UN_OPS = {0: "neg", 1: "abs", 4: "step", 8: "mov"}          # u index -> op (others: C++ switch)


def gen(P, PRIVATE_DISPATCH):
    """asm operands: %0..%(P-1) acc, %P sacc, %(P+1) pc (shared-window address of the instruction word to run
    next; on return: of the instruction the block has no body for), then wbase, sbase, hs (in)."""
    acc = [f"%{k}" for k in range(P)]
    sacc, pc = f"%{P}", f"%{P + 1}"
    wbase, sbase, hs = (f"%{P + 2 + i}" for i in range(3))
    L = []

    def emit(s):
        L.append(s)

    def field(reg, which):
        # 'a' = hi & 0xffff, 'b' = hi >> 16, 'd' = lo >> 16
        if which == "a":
            emit(f"and.b32 {reg}, hi, 0xffff;")
        elif which == "b":
            emit(f"shr.u32 {reg}, hi, 16;")
        else:
            emit(f"shr.u32 {reg}, lo, 16;")

    def wide_addr(reg, which):          # pre-scaled 16-byte units
        field(reg, which)
        emit(f"shl.b32 {reg}, {reg}, 4;")
        emit(f"add.u32 {reg}, {reg}, {wbase};")

    def scal_addr(reg, which):
        field(reg, which)
        emit(f"shl.b32 {reg}, {reg}, 3;")
        emit(f"add.u32 {reg}, {reg}, {sbase};")

    def next_word():
        # All fields of the current word are in registers by now: fetch the next word over it.  Its latency
        # hides behind this body's own loads and arithmetic.
        emit(f"add.u32 {pc}, {pc}, 8;")
        emit(f"ld.shared.v2.b32 {{lo, hi}}, [{pc}];")

    private_sites = []

    def dispatch(private=False):
        if private:
            # a dispatch of its own, with its own copy of the table: the index arithmetic and the table load can
            # then be scheduled early in the body, overlapping the body's own loads and arithmetic
            k = len(private_sites)
            private_sites.append(k)
            emit("and.b32 t, lo, 511;")
            emit(f"TBL{k}: .branchtargets @TARGETS@;")
            emit(f"brx.idx t, TBL{k};")
            return
        # ONE shared dispatch site.  (With an indexed branch at the end of every body ptxas expands each site
        # into its own 512-entry table of BRA instructions -- 1.5 MB of code, measured -- instead of the single
        # constant-bank table + LDC/BRX it builds for one site.)
        emit("bra LOOP;")

    def wide_load(regs, addr):
        if P == 1:
            emit(f"ld.shared.f64 {regs[0]}, [{addr}];")
        else:
            emit(f"ld.shared.v2.f64 {{{regs[0]}, {regs[1]}}}, [{addr}];")
            if P == 4:
                emit(f"add.u32 {addr}, {addr}, {hs};")
                emit(f"ld.shared.v2.f64 {{{regs[2]}, {regs[3]}}}, [{addr}];")

    def wide_store(addr):
        if P == 1:
            emit(f"st.shared.f64 [{addr}], {acc[0]};")
        else:
            emit(f"st.shared.v2.f64 [{addr}], {{{acc[0]}, {acc[1]}}};")
            if P == 4:
                emit(f"add.u32 {addr}, {addr}, {hs};")
                emit(f"st.shared.v2.f64 [{addr}], {{{acc[2]}, {acc[3]}}};")

    def binop(op, d, x, y):
        if op == "add":
            emit(f"add.rn.f64 {d}, {x}, {y};")
        elif op == "mul":
            emit(f"mul.rn.f64 {d}, {x}, {y};")
        else:
            emit(f"setp.{'gt' if op == 'max' else 'lt'}.f64 p, {y}, {x};")
            emit(f"setp.nan.f64 q, {x}, {x};")
            emit("or.pred p, p, q;")
            emit(f"selp.f64 {d}, {y}, {x}, p;")

    def unop(op, d, x):
        if op == "neg":
            emit(f"neg.f64 {d}, {x};")
        elif op == "abs":
            emit(f"abs.f64 {d}, {x};")
        elif op == "step":
            emit(f"setp.ge.f64 p, {x}, 0d0000000000000000;")
            emit(f"selp.f64 {d}, 0d3FF0000000000000, 0d0000000000000000, p;")
        elif d != x:
            emit(f"mov.f64 {d}, {x};")

    xs = [f"x{k}" for k in range(P)]
    ys = [f"y{k}" for k in range(P)]

    def body(label, kinds, store, shape, compute, hot=False):
        """One handler: addresses from the current word, next word, operand loads, arithmetic, store, dispatch."""
        emit(f"{label}:")
        srcs = []
        for which, kind, tmp in kinds:
            if kind == W:
                wide_addr("ad" + which, which)
            elif kind == S:
                scal_addr("ad" + which, which)
        if store:
            (wide_addr if shape == "wide" else scal_addr)("add_", "d")
        next_word()
        for which, kind, tmp in kinds:
            if kind == A:
                srcs.append(acc)
            elif kind == T:
                srcs.append([sacc] * P)
            elif kind == S:
                emit(f"ld.shared.f64 {tmp[0]}, [ad{which}];")
                srcs.append([tmp[0]] * P)
            else:
                wide_load(tmp, "ad" + which)
                srcs.append(tmp)
        compute(srcs)
        if store:
            if shape == "wide":
                wide_store("add_")
            else:
                emit(f"st.shared.f64 [add_], {sacc};")     # the same bits from every lane of the block
        dispatch(private=hot and PRIVATE_DISPATCH)

    targets = ["GEN"] * 512          # everything without a body here, END (0) and YIELD (1) included, leaves the block
    emit("{")
    emit(".reg .u32 t, lo, hi, ada, adb, add_;")
    emit(".reg .f64 " + ", ".join(xs + ys) + ";")
    emit(".reg .pred p, q;")
    emit("TBL: .branchtargets @TARGETS@;")
    emit(f"ld.shared.v2.b32 {{lo, hi}}, [{pc}];")
    emit("LOOP:")
    emit("and.b32 t, lo, 511;")              # handler id + the store flag (bit 8): stores have their own bodies
    emit("brx.idx t, TBL;")

    for store in (0, 1):
        sfx = "S" if store else ""
        # wide binary
        for oi, op in enumerate(BIN_OPS):
            for ka in range(4):
                for kb in range(4):
                    hid = H_BIN + oi * 16 + ka * 4 + kb
                    targets[hid + 256 * store] = f"H{hid}{sfx}"

                    def compute(srcs, op=op, ka=ka, kb=kb):
                        x, y = srcs
                        if ka in (S, T) and kb in (S, T):
                            binop(op, acc[0], x[0], y[0])           # both scalar: one evaluation serves the P pixels
                            for k in range(1, P):
                                emit(f"mov.f64 {acc[k]}, {acc[0]};")
                        else:
                            for k in range(P):
                                binop(op, acc[k], x[k], y[k])
                    body(f"H{hid}{sfx}", [("a", ka, xs), ("b", kb, ys)], store, "wide", compute,
                         hot=(op in ("add", "mul") and (ka, kb) in ((A, S), (A, W), (S, A), (W, A), (W, S), (W, W), (S, W)))
                             or (op in ("max", "min") and (ka, kb) in ((A, W), (W, A), (A, S), (S, A))))
        # wide binary with the accumulator operand negated first (a `neg` fused into its only consumer)
        for oi, op in enumerate(BIN_OPS):
            for c, (ka, kb) in enumerate(((A, W), (A, S), (A, T), (W, A), (S, A), (T, A))):
                hid = H_BINN + oi * 6 + c
                targets[hid + 256 * store] = f"H{hid}{sfx}"

                def compute(srcs, op=op, ka=ka):
                    x, y = srcs
                    neg = xs if ka == A else ys            # the temporaries of the operand that is the accumulator are free
                    for k in range(P):
                        emit(f"neg.f64 {neg[k]}, {acc[k]};")
                    for k in range(P):
                        binop(op, acc[k], neg[k] if ka == A else x[k], y[k] if ka == A else neg[k])
                body(f"H{hid}{sfx}", [("a", ka, xs), ("b", kb, ys)], store, "wide", compute)
        # wide unary
        for u, op in UN_OPS.items():
            for ka in range(4):
                hid = H_UN + u * 4 + ka
                targets[hid + 256 * store] = f"H{hid}{sfx}"

                def compute(srcs, op=op, ka=ka):
                    (x,) = srcs
                    if ka in (S, T):
                        unop(op, acc[0], x[0])
                        for k in range(1, P):
                            emit(f"mov.f64 {acc[k]}, {acc[0]};")
                    else:
                        for k in range(P):
                            unop(op, acc[k], x[k])
                body(f"H{hid}{sfx}", [("a", ka, xs)], store, "wide", compute, hot=(op in ("neg", "step") and ka == A))
        # scalar shape: binary (kinds S/T only: index = (ka == T) * 2 + (kb == T)), unary
        for oi, op in enumerate(BIN_OPS):
            for ta in range(2):
                for tb in range(2):
                    hid = H_SBIN + oi * 4 + ta * 2 + tb
                    targets[hid + 256 * store] = f"H{hid}{sfx}"

                    def compute(srcs, op=op):
                        x, y = srcs
                        binop(op, sacc, x[0], y[0])
                    body(f"H{hid}{sfx}", [("a", T if ta else S, xs), ("b", T if tb else S, ys)], store, "scalar", compute)
        for u, op in UN_OPS.items():
            for ta in range(2):
                hid = H_SUN + u * 2 + ta
                targets[hid + 256 * store] = f"H{hid}{sfx}"

                def compute(srcs, op=op):
                    (x,) = srcs
                    unop(op, sacc, x[0])
                body(f"H{hid}{sfx}", [("a", T if ta else S, xs)], store, "scalar", compute)
    emit("GEN:")                               # pc addresses the instruction that has no body here
    emit("}")
    return [s.replace("@TARGETS@", ", ".join(targets)) for s in L]


def main():
    out = ["// GENERATED by tools/gen_interp_dispatch.py -- do not edit; see that file for the why and the semantics.",
           "// One macro per pixels-per-thread count: MR_INTERP_LOOP_P<P>(acc..., sacc, pc, wbase, sbase, hs)."]
    # (A variant whose hottest bodies end in a dispatch of their own -- gen(P, True) -- was measured slower:
    # chess_1k 104.7 vs 116.6 Mpixel/s; ptxas expands every extra site into its own table of BRA instructions.)
    for P, private in ((1, False), (2, False), (4, False)):
        lines = gen(P, private)
        accs = ", ".join(f"ACC{k}" for k in range(P))
        out.append(f"#define MR_INTERP_LOOP{'X' if private else ''}_P{P}({accs}, SACC, PC, WBASE, SBASE, HS) \\")
        out.append("    asm volatile( \\")
        for s in lines:
            out.append(f'        "{s}\\n" \\')
        cons_out = ", ".join([f'"+d"(ACC{k})' for k in range(P)] + ['"+d"(SACC)', '"+r"(PC)'])
        out.append(f'        : {cons_out} \\')
        out.append('        : "r"(WBASE), "r"(SBASE), "r"(HS) \\')
        out.append('        : "memory")')
        out.append("")
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
