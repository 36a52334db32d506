// SSA program -> accumulator bytecode.  See bytecode.hpp for the format.
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
    BcGen(const Program& p, Bytecode& b) : P(p), B(b) {}
    std::string err;

    void run() {
        const size_t n = P.nodes.size();
        uses_left.assign(n, 0);
        slot.assign(n, -1);
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
        for (int c = 0; c < 3; c++) uses_left[P.root[c]]++;
        // X and Y are preloaded into slots 0 and 1 and stay there.
        for (size_t i = 0; i < n; i++) {
            if (P.nodes[i].op == OP_X) { slot[i] = 0; uses_left[i] = 0x7fffffff; }
            if (P.nodes[i].op == OP_Y) { slot[i] = 1; uses_left[i] = 0x7fffffff; }
        }
        B.n_slots = 2;

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
            Op o = P.nodes[P.root[c]].op;
            if (o == OP_CONST || o == OP_X || o == OP_Y) {
                to_acc(P.root[c]);
                emit(outs[c], 0);
            }
        }
        emit(BC_END, 0);
    }

private:
    const Program& P;
    Bytecode& B;
    std::vector<uint32_t> uses_left;
    std::vector<int32_t> slot, kidx;
    std::vector<uint32_t> free_slots;
    uint32_t acc_holds = NONE;

    bool is_const(uint32_t id) const { return P.nodes[id].op == OP_CONST; }
    bool has_slot(uint32_t id) const { return slot[id] >= 0; }
    bool avail(uint32_t id) const { return is_const(id) || has_slot(id); }

    uint32_t alloc_slot() {
        if (!free_slots.empty()) { uint32_t s = free_slots.back(); free_slots.pop_back(); return s; }
        if (B.n_slots >= 65535) { err = "program needs more than 65535 live values"; return 0; }
        return B.n_slots++;
    }
    void emit(BcOp op, uint32_t operand) { B.code.push_back(bc_encode(op, operand)); }
    // The instruction just emitted produced the accumulator value: make it also store to a slot.
    uint32_t store_last() {
        uint32_t s = alloc_slot();
        B.code.back() |= BC_FLAG_STORE | (uint64_t(s) << 16);
        return s;
    }
    void consume(uint32_t id) {
        if (is_const(id)) return;
        if (uses_left[id] == 0) { err = "internal: value consumed more often than it is used"; return; }
        if (--uses_left[id] == 0 && slot[id] >= 2) { free_slots.push_back(uint32_t(slot[id])); slot[id] = -1; }
    }

    void to_acc(uint32_t id) {
        if (acc_holds == id && !is_const(id)) return;
        if (is_const(id)) { emit(BC_LD_K, uint32_t(kidx[id])); acc_holds = NONE; return; }
        if (!has_slot(id)) { err = "internal: operand is neither in the accumulator nor in a slot"; return; }
        emit(BC_LD_S, uint32_t(slot[id]));
        acc_holds = id;
    }

    // Does `user` read `id`, and can it take it from the accumulator?
    bool next_takes_from_acc(uint32_t id, uint32_t user) const {
        if (user == NONE) return false;
        const Node& u = P.nodes[user];
        if (op_is_unary(u.op)) return u.a == id;
        if (op_is_binary(u.op)) return (u.a == id) != (u.b == id);   // exactly one operand is `id`
        return false;
    }

    void gen(uint32_t id, uint32_t next) {
        const Node& n = P.nodes[id];
        int32_t tmp = -1;
        if (op_is_unary(n.op)) {
            to_acc(n.a);
            if (!err.empty()) return;
            emit(BcOp(BC_NEG + (n.op - OP_NEG)), 0);
            consume(n.a);
        } else {
            uint32_t a = n.a, b = n.b;
            const bool tex = n.op == OP_TEX;
            // the operand that is NOT in the accumulator must be addressable: a slot, or (except for
            // texture coordinates) a constant
            auto addressable = [&](uint32_t v) { return tex ? has_slot(v) : avail(v); };
            bool a_in_acc;
            if (acc_holds == a && !is_const(a) && addressable(b)) a_in_acc = true;
            else if (acc_holds == b && !is_const(b) && addressable(a)) a_in_acc = false;
            else if (is_const(b)) a_in_acc = true;        // load a, then `op constant`
            else if (is_const(a)) a_in_acc = false;       // load b, then `constant op` (reversed form)
            else a_in_acc = true;                         // both in slots: load a, then `op slot`
            uint32_t via_acc = a_in_acc ? a : b, other = a_in_acc ? b : a;
            if (tex && is_const(other)) {                 // a constant coordinate must sit in a slot
                emit(BC_LD_K, uint32_t(kidx[other]));
                acc_holds = NONE;
                tmp = int32_t(store_last());
            }
            uint32_t other_operand;
            bool k = false;
            if (tmp >= 0) other_operand = uint32_t(tmp);
            else if (is_const(other)) { k = true; other_operand = uint32_t(kidx[other]); }
            else if (has_slot(other)) other_operand = uint32_t(slot[other]);
            else { err = "internal: second operand is not addressable"; return; }
            to_acc(via_acc);
            if (!err.empty()) return;
            BcOp op = BC_END;
            switch (n.op) {
            case OP_ADD: op = k ? BC_ADD_K : BC_ADD_S; break;          // commutative: one form
            case OP_MUL: op = k ? BC_MUL_K : BC_MUL_S; break;
            case OP_MAX: op = a_in_acc ? (k ? BC_MAX_K : BC_MAX_S) : (k ? BC_MAXR_K : BC_MAXR_S); break;
            case OP_MIN: op = a_in_acc ? (k ? BC_MIN_K : BC_MIN_S) : (k ? BC_MINR_K : BC_MINR_S); break;
            case OP_TEX:
                if (n.imm > 0xffff) { err = "texture index too large for the interpreter back end"; return; }
                // TEX: x = slot, y = acc.  TEXR: x = acc, y = slot.
                op = a_in_acc ? BC_TEXR_S : BC_TEX_S;
                other_operand |= n.imm << 16;
                break;
            default: err = "internal: unexpected binary op"; return;
            }
            emit(op, other_operand);
            consume(a);
            consume(b);
            if (tmp >= 0) free_slots.push_back(uint32_t(tmp));
        }
        if (!err.empty()) return;
        acc_holds = id;
        // channel outputs are written the moment their value exists
        static const BcOp outs[3] = {BC_OUT_R, BC_OUT_G, BC_OUT_B};
        // A slot is needed unless the only remaining reader is the next instruction via the accumulator.
        uint32_t root_refs = 0;
        for (int c = 0; c < 3; c++) if (P.root[c] == id) root_refs++;
        uint32_t other_uses = uses_left[id] - root_refs;
        if (other_uses > 1 || (other_uses == 1 && !next_takes_from_acc(id, next))) {
            slot[id] = int32_t(store_last());
        }
        for (int c = 0; c < 3; c++) {
            if (P.root[c] == id) { emit(outs[c], 0); consume(id); }
        }
    }
};

struct Job { const Program* p; Bytecode* b; std::string err; };
void job_main(void* arg) {
    Job* j = static_cast<Job*>(arg);
    BcGen g(*j->p, *j->b);
    g.run();
    j->err = g.err;
}

}  // namespace

bool compile_bytecode(const Program& prog, Bytecode* out, std::string* err) {
    out->code.clear();
    out->consts.clear();
    Job j{&prog, out, {}};
    job_main(&j);
    if (!j.err.empty()) { if (err) *err = j.err; return false; }
    if (out->consts.empty()) out->consts.push_back(0.0);
    return true;
}

}  // namespace maray
