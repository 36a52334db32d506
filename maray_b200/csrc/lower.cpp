// Expr tree -> hash-consed SSA program.  See program.hpp for the contract.
#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "program.hpp"

namespace maray {

const char* op_name(Op o) {
    static const char* names[OP_COUNT] = {"const", "x", "y", "neg", "abs", "recip", "sqrt", "step", "sin", "exp",
                                          "ln", "add", "mul", "max", "min", "tex"};
    return o < OP_COUNT ? names[o] : "?";
}

namespace {

struct Key {
    uint32_t op, a, b, imm;
    uint64_t kbits;
    bool operator==(const Key& o) const { return op == o.op && a == o.a && b == o.b && imm == o.imm && kbits == o.kbits; }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        uint64_t h = k.kbits * 0x9e3779b97f4a7c15ull;
        h ^= (uint64_t(k.op) << 56) ^ (uint64_t(k.a) << 28) ^ uint64_t(k.b) ^ (uint64_t(k.imm) << 40);
        h *= 0xff51afd7ed558ccdull;
        return size_t(h ^ (h >> 31));
    }
};

constexpr uint32_t NONE = 0xffffffffu;

struct Binding { const Expr* def; uint32_t value; uint8_t state; };   // state: 0 new, 1 lowering, 2 done

struct Scope {
    const Expr* let;
    std::vector<Binding> vars;
    std::unordered_map<uint64_t, uint32_t> index;   // id -> first definition with that id
    Scope* up;
};

uint64_t bits_of(double v) { uint64_t b; std::memcpy(&b, &v, 8); return b; }

class Lowering {
public:
    Lowering(const std::vector<TextureDim>& tex) : tex_(tex) {}

    std::vector<Node> nodes;
    std::string err;

    uint32_t konst(double v) {
        Key k{OP_CONST, 0, 0, 0, bits_of(v)};
        auto it = table_.find(k);
        if (it != table_.end()) return it->second;
        Node n{OP_CONST, DEP_CONST, 0, 0, 0, v};
        nodes.push_back(n);
        uint32_t id = uint32_t(nodes.size() - 1);
        table_.emplace(k, id);
        return id;
    }

    uint32_t leaf(Op op) {
        Key k{op, 0, 0, 0, 0};
        auto it = table_.find(k);
        if (it != table_.end()) return it->second;
        Node n{op, uint8_t(op == OP_X ? DEP_X : DEP_Y), 0, 0, 0, 0.0};
        nodes.push_back(n);
        uint32_t id = uint32_t(nodes.size() - 1);
        table_.emplace(k, id);
        return id;
    }

    uint32_t unary(Op op, uint32_t a) {
        if (nodes[a].op == OP_CONST) {
            // Host folding: the same IEEE operation, and for sin/exp/ln the host libm -- which is
            // what the reference itself calls (reference src/lib.rs:648-650).
            double v = nodes[a].k, r = 0.0;
            switch (op) {
            case OP_NEG: r = -v; break;
            case OP_ABS: r = std::fabs(v); break;
            case OP_RECIP: r = 1.0 / v; break;
            case OP_SQRT: r = std::sqrt(v); break;
            case OP_STEP: r = sem_step(v); break;
            case OP_SIN: r = std::sin(v); break;
            case OP_EXP: r = std::exp(v); break;
            case OP_LN: r = std::log(v); break;
            default: break;
            }
            return konst(r);
        }
        return intern(op, a, 0, 0);
    }

    uint32_t binary(Op op, uint32_t a, uint32_t b) {
        if (nodes[a].op == OP_CONST && nodes[b].op == OP_CONST) {
            double x = nodes[a].k, y = nodes[b].k, r = 0.0;
            switch (op) {
            case OP_ADD: r = x + y; break;
            case OP_MUL: r = x * y; break;
            case OP_MAX: r = sem_max(x, y); break;
            case OP_MIN: r = sem_min(x, y); break;
            default: break;
            }
            return konst(r);
        }
        return intern(op, a, b, 0);
    }

    // App(id, a, b) with the default texture runtime (reference src/textures.rs:14-65).
    uint32_t app(uint32_t id, uint32_t a, uint32_t b) {
        uint32_t img = id / 5, k = id % 5;
        if (img >= tex_.size()) {
            err = "App id " + std::to_string(id) + " is outside the texture runtime's function table (" +
                  std::to_string(tex_.size() * 5) + " functions for " + std::to_string(tex_.size()) + " textures)";
            return NONE;
        }
        if (k == 3) return konst(double(tex_[img].w));    // fun_image_width  (arguments ignored)
        if (k == 4) return konst(double(tex_[img].h));    // fun_image_height
        return intern(OP_TEX, a, b, img * 4 + k);
    }

    uint32_t lower(const Expr* e, Scope* sc) {
        if (!err.empty()) return NONE;
        switch (e->tag) {
        case T_ARC: case T_DECOR: return lower(e->a, sc);
        case T_X: return leaf(OP_X);
        case T_Y: return leaf(OP_Y);
        case T_TAU: return konst(6.283185307179586);
        case T_E: return konst(2.718281828459045);
        case T_NAT: return konst(double(e->n));
        case T_VAR: return var(e->n, sc);
        case T_NEG: case T_ABS: case T_RECIP: case T_SQRT:
        case T_STEP: case T_SIN: case T_EXP: case T_LN: {
            uint32_t a = lower(e->a, sc);
            if (a == NONE) return NONE;
            return unary(Op(OP_NEG + (e->tag - T_NEG)), a);
        }
        case T_ADD: case T_MUL: case T_MAX: case T_MIN: {
            uint32_t a = lower(e->a, sc);
            uint32_t b = lower(e->b, sc);
            if (a == NONE || b == NONE) return NONE;
            return binary(Op(OP_ADD + (e->tag - T_ADD)), a, b);
        }
        case T_APP: {
            uint32_t a = lower(e->a, sc);
            uint32_t b = lower(e->b, sc);
            if (a == NONE || b == NONE) return NONE;
            return app(e->app_id, a, b);
        }
        case T_LET: {
            Scope in;
            in.let = e; in.up = sc;
            in.vars.resize(e->n_vars);
            for (uint64_t i = 0; i < e->n_vars; i++) {
                in.vars[i] = Binding{e->vars[i].def, NONE, 0};
                in.index.emplace(e->vars[i].id, uint32_t(i));   // keeps the FIRST definition of an id
            }
            // Lower the definitions in file order (keeps recursion shallow for the usual
            // dependency-ordered contexts); forward references are resolved on demand by var().
            for (uint64_t i = 0; i < e->n_vars; i++) {
                auto first = in.index.find(e->vars[i].id);
                if (first->second != i) continue;               // shadowed duplicate id: never visible
                if (force(&in, uint32_t(i)) == NONE) return NONE;
            }
            return lower(e->a, &in);
        }
        default: break;
        }
        err = "unsupported expression variant";
        return NONE;
    }

private:
    uint32_t intern(Op op, uint32_t a, uint32_t b, uint32_t imm) {
        Key k{op, a, b, imm, 0};
        auto it = table_.find(k);
        if (it != table_.end()) return it->second;
        uint8_t dep = uint8_t(nodes[a].dep | (op_is_binary(op) ? nodes[b].dep : 0));
        Node n{op, dep, a, op_is_binary(op) ? b : 0, imm, 0.0};
        nodes.push_back(n);
        uint32_t id = uint32_t(nodes.size() - 1);
        table_.emplace(k, id);
        return id;
    }

    uint32_t force(Scope* s, uint32_t i) {
        Binding& bnd = s->vars[i];
        if (bnd.state == 2) return bnd.value;
        if (bnd.state == 1) {
            err = "cyclic Let definitions (variable $" + std::to_string(s->let->vars[i].id) + ")";
            return NONE;
        }
        bnd.state = 1;
        uint32_t v = lower(bnd.def, s);
        s->vars[i].value = v;
        s->vars[i].state = 2;
        return v;
    }

    uint32_t var(uint64_t name, Scope* sc) {
        for (Scope* s = sc; s; s = s->up) {
            auto it = s->index.find(name);
            if (it != s->index.end()) return force(s, it->second);
        }
        err = "unbound variable $" + std::to_string(name);
        return NONE;
    }

    const std::vector<TextureDim>& tex_;
    std::unordered_map<Key, uint32_t, KeyHash> table_;
};

struct LowerJob {
    const Scene* scene;
    const std::vector<TextureDim>* tex;
    Program* out;
    std::string* err;
    bool ok;
};

void lower_job(void* arg) {
    LowerJob* j = static_cast<LowerJob*>(arg);
    Lowering L(*j->tex);
    uint32_t roots[3];
    for (int c = 0; c < 3; c++) {
        roots[c] = L.lower(j->scene->color[c], nullptr);
        if (roots[c] == NONE) { *j->err = L.err.empty() ? "lowering failed" : L.err; j->ok = false; return; }
    }
    const std::vector<Node>& all = L.nodes;   // creation order is already topological

    // Register-need estimate (Sethi-Ullman on the tree view) to pick operand evaluation order:
    // visiting the needier operand first keeps fewer values live.  Order does not affect values.
    std::vector<uint32_t> need(all.size(), 0);
    for (size_t i = 0; i < all.size(); i++) {
        const Node& n = all[i];
        if (n.op == OP_CONST) need[i] = 0;
        else if (n.op == OP_X || n.op == OP_Y) need[i] = 1;
        else if (op_is_unary(n.op)) need[i] = std::max(need[n.a], 1u);
        else {
            uint32_t na = need[n.a], nb = need[n.b];
            need[i] = std::max(1u, na == nb ? na + 1 : std::max(na, nb));
        }
    }

    // Depth-first post-order from R, G, B: prunes unreachable values and fixes the schedule.
    std::vector<uint32_t> remap(all.size(), NONE);
    std::vector<Node> out_nodes;
    out_nodes.reserve(all.size());
    struct Frame { uint32_t id; uint8_t stage; };
    std::vector<Frame> stack;
    for (int c = 0; c < 3; c++) {
        stack.push_back({roots[c], 0});
        while (!stack.empty()) {
            Frame& f = stack.back();
            uint32_t id = f.id;
            if (remap[id] != NONE) { stack.pop_back(); continue; }
            const Node& n = all[id];
            bool un = op_is_unary(n.op), bin = op_is_binary(n.op);
            uint32_t first = n.a, second = n.b;
            if (bin && need[n.b] > need[n.a]) std::swap(first, second);
            if (f.stage == 0) {
                f.stage = 1;
                if (un || bin) { stack.push_back({first, 0}); continue; }
            }
            if (f.stage == 1) {
                f.stage = 2;
                if (bin) { stack.push_back({second, 0}); continue; }
            }
            Node m = n;
            if (un || bin) m.a = remap[n.a];
            if (bin) m.b = remap[n.b];
            out_nodes.push_back(m);
            remap[id] = uint32_t(out_nodes.size() - 1);
            stack.pop_back();
        }
    }

    Program* P = j->out;
    P->nodes.swap(out_nodes);
    for (int c = 0; c < 3; c++) P->root[c] = remap[roots[c]];
    P->n_textures = uint32_t(j->tex->size());
    P->order.clear();
    ProgramStats st;
    st.tree_nodes = j->scene->tree_nodes[0] + j->scene->tree_nodes[1] + j->scene->tree_nodes[2];
    st.dag_nodes = P->nodes.size();
    std::vector<uint32_t> depth(P->nodes.size(), 0);
    for (size_t i = 0; i < P->nodes.size(); i++) {
        const Node& n = P->nodes[i];
        if (n.op == OP_CONST) { st.n_const++; continue; }
        P->order.push_back(uint32_t(i));
        st.op_count[n.op]++;
        switch (n.dep) {
        case DEP_X: st.n_x++; break;
        case DEP_Y: st.n_y++; break;
        default: st.n_xy++; break;       // includes texture fetches at constant coordinates
        }
        uint32_t d = 0;
        if (op_is_unary(n.op) || op_is_binary(n.op)) d = depth[n.a];
        if (op_is_binary(n.op)) d = std::max(d, depth[n.b]);
        depth[i] = d + 1;
        st.depth = std::max(st.depth, depth[i]);
    }
    P->stats = st;
    j->ok = true;
}

}  // namespace

bool lower_scene(const Scene& scene, const std::vector<TextureDim>& textures, Program* out, std::string* err) {
    std::string local;
    LowerJob j{&scene, &textures, out, err ? err : &local, false};
    try {
        run_with_big_stack(lower_job, &j);
    } catch (const std::bad_alloc&) {
        if (err) *err = "out of memory while lowering";
        return false;
    }
    return j.ok;
}

}  // namespace maray
