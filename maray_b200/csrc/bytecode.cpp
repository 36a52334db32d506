// SSA program -> accumulator bytecode.  See bytecode.hpp for the format.
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
        computed.assign(n, 0);
        kidx.assign(n, -1);
        need.assign(n, 0);
        for (size_t i = 0; i < n; i++) {
            const Node& nd = P.nodes[i];
            if (nd.op == OP_CONST) {
                kidx[i] = int32_t(B.consts.size());
                B.consts.push_back(nd.k);
                continue;
            }
            if (nd.op == OP_X || nd.op == OP_Y) { need[i] = 1; continue; }
            uses_left[nd.a]++;
            if (op_is_binary(nd.op)) uses_left[nd.b]++;
            if (op_is_unary(nd.op)) need[i] = std::max(need[nd.a], 1u);
            else {
                uint32_t na = need[nd.a], nb = need[nd.b];
                need[i] = std::max(1u, na == nb ? na + 1 : std::max(na, nb));
            }
        }
        for (int c = 0; c < 3; c++) uses_left[P.root[c]]++;
        // X and Y are preloaded into slots 0 and 1 and stay there.
        for (size_t i = 0; i < n; i++) {
            if (P.nodes[i].op == OP_X) { slot[i] = 0; computed[i] = 1; uses_left[i] = 0x7fffffff; }
            if (P.nodes[i].op == OP_Y) { slot[i] = 1; computed[i] = 1; uses_left[i] = 0x7fffffff; }
        }
        B.n_slots = 2;
        static const BcOp outs[3] = {BC_OUT_R, BC_OUT_G, BC_OUT_B};
        for (int c = 0; c < 3 && err.empty(); c++) {
            to_acc(P.root[c]);
            emit(outs[c], 0);
        }
        emit(BC_END, 0);
    }

private:
    const Program& P;
    Bytecode& B;
    std::vector<uint32_t> uses_left, need;
    std::vector<int32_t> slot, kidx;
    std::vector<uint8_t> computed;
    std::vector<uint32_t> free_slots;
    uint32_t acc_holds = NONE;

    bool is_const(uint32_t id) const { return P.nodes[id].op == OP_CONST; }
    bool has_slot(uint32_t id) const { return slot[id] >= 0; }

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
        if (--uses_left[id] == 0 && slot[id] >= 2) { free_slots.push_back(uint32_t(slot[id])); slot[id] = -1; }
    }

    void to_acc(uint32_t id) {
        if (!err.empty()) return;
        if (is_const(id)) { emit(BC_LD_K, uint32_t(kidx[id])); acc_holds = NONE; return; }
        if (acc_holds == id) { consume(id); return; }
        if (computed[id]) {
            if (!has_slot(id)) { err = "internal: value consumed after it was dropped"; return; }
            emit(BC_LD_S, uint32_t(slot[id]));
            acc_holds = id;
            consume(id);
            return;
        }
        gen(id);
        consume(id);
    }

    // Computes `id` (if needed) and makes sure it sits in a slot.
    void to_slot(uint32_t id) {
        if (!computed[id]) gen(id);
        if (!err.empty() || has_slot(id)) return;
        if (acc_holds != id) { err = "internal: value not in accumulator"; return; }
        slot[id] = int32_t(store_last());
    }

    void gen(uint32_t id) {
        if (!err.empty()) return;
        const Node& n = P.nodes[id];
        if (op_is_unary(n.op)) {
            to_acc(n.a);
            emit(BcOp(BC_NEG + (n.op - OP_NEG)), 0);
        } else if (n.op == OP_TEX) {
            gen_tex(n);
        } else {
            uint32_t a = n.a, b = n.b;
            auto avail = [&](uint32_t v) { return is_const(v) || has_slot(v); };
            if (!avail(a) && !avail(b)) to_slot(need[b] > need[a] ? b : a);
            if (!err.empty()) return;
            bool a_in_acc;
            if (avail(a) && avail(b)) a_in_acc = !(acc_holds == b && acc_holds != a);
            else a_in_acc = avail(b);
            uint32_t via_acc = a_in_acc ? a : b, other = a_in_acc ? b : a;
            // Pin `other` while the accumulator operand is produced (its slot must survive).
            to_acc(via_acc);
            if (!err.empty()) return;
            bool k = is_const(other);
            uint32_t operand = k ? uint32_t(kidx[other]) : uint32_t(slot[other]);
            BcOp op = BC_END;
            switch (n.op) {
            case OP_ADD: op = k ? BC_ADD_K : BC_ADD_S; break;          // commutative: one form
            case OP_MUL: op = k ? BC_MUL_K : BC_MUL_S; break;
            case OP_MAX: op = a_in_acc ? (k ? BC_MAX_K : BC_MAX_S) : (k ? BC_MAXR_K : BC_MAXR_S); break;
            case OP_MIN: op = a_in_acc ? (k ? BC_MIN_K : BC_MIN_S) : (k ? BC_MINR_K : BC_MINR_S); break;
            default: err = "internal: unexpected binary op"; return;
            }
            emit(op, operand);
            consume(other);
        }
        computed[id] = 1;
        acc_holds = id;
        if (uses_left[id] > 1) slot[id] = int32_t(store_last());
    }

    // App: both coordinates are values; the one not in the accumulator must be in a slot
    // (a constant coordinate is first materialised into a temporary slot).
    void gen_tex(const Node& n) {
        uint32_t a = n.a, b = n.b;
        int32_t tmp = -1;
        auto slot_of_const = [&](uint32_t v) {
            emit(BC_LD_K, uint32_t(kidx[v]));
            acc_holds = NONE;
            tmp = int32_t(store_last());
            return uint32_t(tmp);
        };
        uint32_t other_slot;
        bool a_in_acc;
        if (is_const(b)) { other_slot = slot_of_const(b); a_in_acc = true; to_acc(a); }
        else if (is_const(a)) { other_slot = slot_of_const(a); a_in_acc = false; to_acc(b); }
        else {
            if (!has_slot(a) && !has_slot(b)) to_slot(need[b] > need[a] ? b : a);
            if (!err.empty()) return;
            if (has_slot(a) && has_slot(b)) a_in_acc = !(acc_holds == b && acc_holds != a);
            else a_in_acc = has_slot(b);
            uint32_t via_acc = a_in_acc ? a : b, other = a_in_acc ? b : a;
            to_acc(via_acc);
            if (!err.empty()) return;
            other_slot = uint32_t(slot[other]);
            consume(other);
        }
        if (!err.empty()) return;
        if (n.imm > 0xffff) { err = "texture index too large for the interpreter back end"; return; }
        // TEX: x = slot, y = acc.  TEXR: x = acc, y = slot.
        emit(a_in_acc ? BC_TEXR_S : BC_TEX_S, (n.imm << 16) | other_slot);
        if (tmp >= 0) free_slots.push_back(uint32_t(tmp));
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
    run_with_big_stack(job_main, &j);
    if (!j.err.empty()) { if (err) *err = j.err; return false; }
    if (out->consts.empty()) out->consts.push_back(0.0);
    return true;
}

}  // namespace maray
