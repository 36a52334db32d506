// SSA program -> accumulator bytecode (version 2).  See bytecode.hpp for the format.
//
// One linear pass over the program's schedule.  The running value stays in the accumulator; a value
// is written to a slot only if somebody other than the very next instruction needs it.  Slots are
// recycled as soon as their last reader has been emitted.
#include "bytecode.hpp"

#include <algorithm>

namespace maray {
namespace {

constexpr uint32_t NONE = 0xffffffffu;

class BcGen {
public:
    BcGen(const Program& p, Bytecode& b, bool row_uniform) : P(p), B(b), uniform(row_uniform) {}
    std::string err;

    void run() {
        const size_t n = P.nodes.size();
        uses_left.assign(n, 0);
        slot.assign(n, -1);
        uslot.assign(n, -1);
        kidx.assign(n, -1);
        for (size_t i = 0; i < n; i++) {
            const Node& nd = P.nodes[i];
            if (nd.op == OP_CONST) {
                kidx[i] = int32_t(B.consts.size());
                B.consts.push_back(nd.k);
                continue;
            }
            if (op_is_unary(nd.op) || op_is_binary(nd.op)) uses_left[nd.a]++;
            if (op_is_binary(nd.op)) uses_left[nd.b]++;
        }
        if (B.consts.size() > 0xffff) { err = "program has more than 65535 distinct constants"; return; }
        for (int c = 0; c < 3; c++) uses_left[P.root[c]]++;
        // X and Y are preloaded into slots 0 and 1 and stay there.
        for (size_t i = 0; i < n; i++) {
            if (P.nodes[i].op == OP_X) { slot[i] = 0; uses_left[i] = 0x7fffffff; }
            if (P.nodes[i].op == OP_Y) { slot[i] = 1; uses_left[i] = 0x7fffffff; }
        }
        B.n_slots = 2;
        B.n_uniform = 0;

        const std::vector<uint32_t>& order = P.order;
        for (size_t i = 0; i < order.size() && err.empty(); i++) {
            uint32_t id = order[i];
            const Node& nd = P.nodes[id];
            if (nd.op == OP_X || nd.op == OP_Y) continue;
            uint32_t next = NONE;
            for (size_t j = i + 1; j < order.size(); j++) {
                Op o = P.nodes[order[j]].op;
                if (o != OP_X && o != OP_Y) { next = order[j]; break; }
            }
            gen(id, next);
        }
        // channels whose value is a constant, X or Y (everything else was written when computed)
        static const BcOp outs[3] = {BC_OUT_R, BC_OUT_G, BC_OUT_B};
        for (int c = 0; c < 3 && err.empty(); c++) {
            uint32_t r = P.root[c];
            Op o = P.nodes[r].op;
            if (o == OP_CONST) emit(outs[c], BC_F_SWAP | BC_F_B_CONST, 0, 0, uint32_t(kidx[r]));
            else if (o == OP_X || o == OP_Y) emit(outs[c], 0, 0, uint32_t(slot[r]), 0);
        }
        emit(BC_END, 0, 0, 0, 0);
    }

private:
    const Program& P;
    Bytecode& B;
    std::vector<uint32_t> uses_left;
    const bool uniform;              // y-only values go to row-uniform slots
    std::vector<int32_t> slot, uslot, kidx;
    std::vector<uint32_t> free_slots;
    uint32_t acc_holds = NONE;
    int32_t last_stored = -1;        // slot written by the previous instruction, or -1
    int32_t last_stored_uni = -1;    // uniform slot written by the previous instruction, or -1

    bool is_const(uint32_t id) const { return P.nodes[id].op == OP_CONST; }
    bool has_slot(uint32_t id) const { return slot[id] >= 0; }
    bool has_uslot(uint32_t id) const { return uslot[id] >= 0; }
    // y-only values (not Y itself, which is preloaded per pixel) are row-uniform
    bool row_uniform_value(uint32_t id) const { return uniform && P.nodes[id].dep == DEP_Y && P.nodes[id].op != OP_Y; }

    uint32_t alloc_slot() {
        if (!free_slots.empty()) { uint32_t s = free_slots.back(); free_slots.pop_back(); return s; }
        if (B.n_slots >= 65535) { err = "program needs more than 65535 live values"; return 0; }
        return B.n_slots++;
    }
    void emit(BcOp op, uint32_t flags, uint32_t dst, uint32_t a, uint32_t b) {
        B.code.push_back(bc_encode(op, flags, dst, a, b));
        last_stored = -1;
        last_stored_uni = -1;
    }
    // The instruction just emitted produced the accumulator value: make it also store to a slot.
    uint32_t store_last() {
        uint32_t s = alloc_slot();
        B.code.back() |= (uint64_t(BC_F_STORE) << 8) | (uint64_t(s) << 16);
        last_stored = int32_t(s);
        return s;
    }
    // Same, to a fresh row-uniform slot (never recycled: other warps of the block may still read it).
    uint32_t store_last_uniform() {
        if (B.n_uniform >= 65535) { err = "program needs more than 65535 row-uniform values"; return 0; }
        uint32_t s = B.n_uniform++;
        B.code.back() |= (uint64_t(BC_F_STORE | BC_F_ST_UNI) << 8) | (uint64_t(s) << 16);
        last_stored_uni = int32_t(s);
        return s;
    }
    void consume(uint32_t id) {
        if (is_const(id)) return;
        if (uses_left[id] == 0) { err = "internal: value consumed more often than it is used"; return; }
        if (--uses_left[id] == 0 && slot[id] >= 2) { free_slots.push_back(uint32_t(slot[id])); slot[id] = -1; }
    }

    // Does `user` read `id`, and can it take it from the accumulator?
    bool next_takes_from_acc(uint32_t id, uint32_t user) const {
        if (user == NONE) return false;
        const Node& u = P.nodes[user];
        if (op_is_unary(u.op)) return u.a == id;
        if (op_is_binary(u.op)) return (u.a == id) != (u.b == id);   // exactly one operand is `id`
        return false;
    }

    // Operand fields for `first` (a value that is in the accumulator or in a slot).
    bool first_fields(uint32_t v, uint32_t* flags, uint32_t* a) {
        if (acc_holds == v && !is_const(v)) { *flags |= BC_F_ACC_A; *a = 0; return true; }
        if (has_slot(v)) { *a = uint32_t(slot[v]); return true; }
        if (has_uslot(v)) { *flags |= BC_F_A_UNI; *a = uint32_t(uslot[v]); return true; }
        err = "internal: first operand is neither in the accumulator nor in a slot";
        return false;
    }
    // Operand fields for `second` (slot, constant, or the value the previous instruction stored).
    bool second_fields(uint32_t v, uint32_t* flags, uint32_t* b) {
        if (is_const(v)) { *flags |= BC_F_B_CONST; *b = uint32_t(kidx[v]); return true; }
        if (has_uslot(v)) {
            *flags |= BC_F_B_UNI;
            *b = uint32_t(uslot[v]);
            if (uslot[v] == last_stored_uni) *flags |= BC_F_FWD_B;
            return true;
        }
        if (!has_slot(v)) { err = "internal: second operand is not in a slot"; return false; }
        *b = uint32_t(slot[v]);
        if (slot[v] == last_stored) *flags |= BC_F_FWD_B;
        return true;
    }

    void gen(uint32_t id, uint32_t next) {
        const Node& n = P.nodes[id];
        uint32_t flags = 0, a = 0, b = 0;
        if (op_is_unary(n.op)) {
            if (!first_fields(n.a, &flags, &a)) return;
            emit(BcOp(BC_NEG + (n.op - OP_NEG)), flags, 0, a, 0);
            consume(n.a);
        } else {
            // first = the operand taken from the accumulator (if any), else a slot operand;
            // SWAP when `first` is the node's second argument.
            uint32_t first = n.a, second = n.b;
            bool swap = false;
            if (acc_holds == n.a && !is_const(n.a)) { /* keep */ }
            else if (acc_holds == n.b && !is_const(n.b)) { first = n.b; second = n.a; swap = true; }
            else if (is_const(n.a)) { first = n.b; second = n.a; swap = true; }   // a constant can only be `second`
            if (is_const(first)) {
                // both operands constant: only App reaches here (arithmetic was folded).  One of them goes
                // through the accumulator -- not through a slot: the kernel fetches the operands of the
                // next instruction before this one stores, so a slot written here would be read stale.
                emit(BC_MOV, BC_F_SWAP | BC_F_B_CONST, 0, 0, uint32_t(kidx[first]));
                acc_holds = NONE;
                flags |= BC_F_ACC_A;
            } else if (!first_fields(first, &flags, &a)) return;
            if (swap) flags |= BC_F_SWAP;
            if (!second_fields(second, &flags, &b)) return;
            BcOp op = BC_END;
            uint32_t dst = 0;
            switch (n.op) {
            case OP_ADD: op = BC_ADD; break;
            case OP_MUL: op = BC_MUL; break;
            case OP_MAX: op = BC_MAX; break;
            case OP_MIN: op = BC_MIN; break;
            case OP_TEX:
                if (n.imm > 0xffff) { err = "texture index too large for the interpreter back end"; return; }
                op = BC_TEX; dst = n.imm;
                break;
            default: err = "internal: unexpected binary op"; return;
            }
            emit(op, flags, dst, a, b);
            consume(n.a);
            consume(n.b);
        }
        if (!err.empty()) return;
        acc_holds = id;
        static const BcOp outs[3] = {BC_OUT_R, BC_OUT_G, BC_OUT_B};
        uint32_t root_refs = 0;
        for (int c = 0; c < 3; c++) if (P.root[c] == id) root_refs++;
        uint32_t other_uses = uses_left[id] - root_refs;
        // A slot is needed unless the only remaining reader is the next instruction via the accumulator.
        if (other_uses > 1 || (other_uses == 1 && !next_takes_from_acc(id, next))) {
            if (n.op == OP_TEX) emit(BC_MOV, BC_F_ACC_A, 0, 0, 0);   // TEX uses the dst field for its texture id
            if (row_uniform_value(id)) uslot[id] = int32_t(store_last_uniform());
            else slot[id] = int32_t(store_last());
        }
        // channel outputs are written the moment their value exists
        for (int c = 0; c < 3; c++) {
            if (P.root[c] == id) {
                emit(outs[c], BC_F_ACC_A, 0, 0, 0);
                consume(id);
            }
        }
    }
};

}  // namespace

bool compile_bytecode(const Program& prog, Bytecode* out, std::string* err, bool row_uniform) {
    out->code.clear();
    out->consts.clear();
    BcGen g(prog, *out, row_uniform);
    g.run();
    if (!g.err.empty()) { if (err) *err = g.err; return false; }
    if (out->consts.empty()) out->consts.push_back(0.0);
    return true;
}

}  // namespace maray
