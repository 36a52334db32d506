// SSA program -> CUDA C++ source for NVRTC.  See codegen.hpp.
#include "codegen.hpp"

#include <algorithm>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace maray {

const char* const kJitKernelName = "maray_jit";

namespace {

// The text of device_sem.cuh, embedded by the build (csrc/Makefile: device_sem_text.inc).
const char* const kDeviceSemText =
#include "device_sem_text.inc"
    ;

void append_const(std::string& s, double v) {
    char buf[64];
    uint64_t bits;
    std::memcpy(&bits, &v, 8);
    if (v != v || v - v != 0.0) {
        // NaN / infinity: no literal form; rebuild from the bit pattern.
        std::snprintf(buf, sizeof buf, "__longlong_as_double(0x%016" PRIx64 "LL)", bits);
    } else {
        std::snprintf(buf, sizeof buf, "%a", v);   // hex float: exact
        if (buf[0] == '-') { s += "(" ; s += buf; s += ")"; return; }
    }
    s += buf;
}

constexpr uint32_t kScratchRows = 8;   // argument rows of the batched sin/exp/ln scratch (as many result rows follow)

struct Emitter {
    const Program& P;
    bool inline_trans;
    const std::vector<int32_t>* bank_index;   // value id -> index in the __constant__ table, or -1
    // In the per-pixel kernel a hoisted value is a load from its table: 1 = column table, 2 = row table.
    const std::vector<uint8_t>* load_kind = nullptr;
    const std::vector<uint32_t>* table_index = nullptr;
    bool scratch_batches = false;
    std::string helper_suffix;      // which instance of the batch helpers this function calls ("" or "_sK")
    // Values that are exactly +0.0 or 1.0 at every pixel (see find_booleans): kind 1 = boolean,
    // kind 2 = the NOT pattern `1 + -(b)`.  Emitted as `bool` logic plus a double shadow.
    const std::vector<uint8_t>* boolean = nullptr;

    bool is_bool(uint32_t id) const { return boolean && (*boolean)[id]; }
    void bool_operand(std::string& s, uint32_t id) const {
        const Node& n = P.nodes[id];
        if (n.op == OP_CONST) { s += (n.k != 0.0) ? "true" : "false"; return; }
        char buf[16];
        std::snprintf(buf, sizeof buf, "b%u", id);
        s += buf;
    }
    // `const bool bN = vN != 0.0;` after a boolean value arrived as a double (frame import, table load)
    void bool_from_double(std::string& s, uint32_t id) const {
        if (!is_bool(id)) return;
        char buf[64];
        std::snprintf(buf, sizeof buf, "  const bool b%u = v%u != 0.0;\n", id, id);
        s += buf;
    }

    void operand(std::string& s, uint32_t id) const {
        const Node& n = P.nodes[id];
        if (n.op == OP_CONST) {
            if (bank_index && (*bank_index)[id] >= 0) {
                char kb[24];
                std::snprintf(kb, sizeof kb, "MRK(%d)", (*bank_index)[id]);
                s += kb;
            } else append_const(s, n.k);
            return;
        }
        if (n.op == OP_X) { s += "X"; return; }
        if (n.op == OP_Y) { s += "Y"; return; }
        char buf[16];
        std::snprintf(buf, sizeof buf, "v%u", id);
        s += buf;
    }

    // Emits order[i..j) -- one batch of same-kind transcendentals -- as x4 / x2 / single out-of-line
    // calls.  The four (two) bodies inside a call are independent, which is where the FP64 pipe gets
    // its instruction-level parallelism in transcendental-heavy programs.
    void batch_calls(std::string& s, const std::vector<uint32_t>& order, size_t i, size_t j) const {
        const char* fn = P.nodes[order[i]].op == OP_SIN ? "sin" : (P.nodes[order[i]].op == OP_EXP ? "exp" : "log");
        char buf[96];
        if (scratch_batches) {
            // arguments -> this thread's scratch rows, one leaf call for the whole batch, results back
            // (rows: kScratchRows; lower.cpp's kBatchMax never exceeds it)
            while (i < j) {
                const size_t k = std::min<size_t>(j - i, kScratchRows);
                if (k == 1) { statement(s, order[i]); i++; continue; }
                for (size_t m = 0; m < k; m++) {
                    std::snprintf(buf, sizeof buf, "  MR_S(%zu) = ", m);
                    s += buf; operand(s, P.nodes[order[i + m]].a); s += ";\n";
                }
                std::snprintf(buf, sizeof buf, "  if (mr_%s_batch%s(%zuu)) mr_%s_fix%s(%zuu);\n", fn, helper_suffix.c_str(), k, fn,
                              helper_suffix.c_str(), k);
                s += buf;
                for (size_t m = 0; m < k; m++) {
                    std::snprintf(buf, sizeof buf, "  const double v%u = MR_R(%zu);\n", order[i + m], m);
                    s += buf;
                }
                i += k;
            }
            return;
        }
        while (i < j) {
            size_t k = (j - i >= 4) ? 4 : ((j - i >= 2) ? 2 : 1);
            if (k == 1) { statement(s, order[i]); i++; continue; }
            std::snprintf(buf, sizeof buf, "  const MrD%zu b%u = mr_%s_x%zu(", k, order[i], fn, k);
            s += buf;
            for (size_t m = 0; m < k; m++) { if (m) s += ", "; operand(s, P.nodes[order[i + m]].a); }
            s += ");\n";
            static const char* fld[4] = {"a", "b", "c", "d"};
            for (size_t m = 0; m < k; m++) {
                std::snprintf(buf, sizeof buf, "  const double v%u = b%u.%s;\n", order[i + m], order[i], fld[m]);
                s += buf;
            }
            i += k;
        }
    }

    void statement(std::string& s, uint32_t id) const {
        const Node& n = P.nodes[id];
        if (n.op == OP_X || n.op == OP_Y) return;   // kernel arguments
        char buf[160];
        if (load_kind && (*load_kind)[id]) {
            if ((*load_kind)[id] == 1) std::snprintf(buf, sizeof buf, "  const double v%u = __ldg(CV + %uu * CW);\n", id, (*table_index)[id]);
            else std::snprintf(buf, sizeof buf, "  const double v%u = __ldg(RV + %uu * RW);\n", id, (*table_index)[id]);
            s += buf;
            bool_from_double(s, id);
            return;
        }
        if (is_bool(id)) {
            // 0/1-valued: compares and bit logic instead of FP64 multiplies, min/max and selects.  The
            // double shadow is what every non-boolean consumer reads; unused shadows are dead code.
            std::snprintf(buf, sizeof buf, "  const bool b%u = ", id);
            s += buf;
            switch (n.op) {
            case OP_STEP: s += "("; operand(s, n.a); s += " >= 0.0)"; break;          // NaN -> false, -0.0 -> true
            case OP_ADD: {                                                           // 1 + -(b)  ==  !b
                const uint32_t neg = P.nodes[n.a].op == OP_NEG ? n.a : n.b;
                s += "!"; bool_operand(s, P.nodes[neg].a);
            } break;
            case OP_MUL: case OP_MIN: s += "("; bool_operand(s, n.a); s += " && "; bool_operand(s, n.b); s += ")"; break;
            case OP_MAX: s += "("; bool_operand(s, n.a); s += " || "; bool_operand(s, n.b); s += ")"; break;
            default: s += "false"; break;
            }
            std::snprintf(buf, sizeof buf, ";\n  const double v%u = b%u ? 1.0 : 0.0;\n", id, id);
            s += buf;
            return;
        }
        std::snprintf(buf, sizeof buf, "  const double v%u = ", id);
        s += buf;
        auto un = [&](const char* pre, const char* post) { s += pre; operand(s, n.a); s += post; };
        auto bin = [&](const char* pre, const char* mid, const char* post) {
            s += pre; operand(s, n.a); s += mid; operand(s, n.b); s += post;
        };
        switch (n.op) {
        case OP_NEG: un("-", ""); break;
        case OP_ABS: un("fabs(", ")"); break;
        case OP_RECIP: un("mr_recip(", ")"); break;
        case OP_SQRT: un("mr_sqrt(", ")"); break;
        case OP_STEP: un("mr_step(", ")"); break;
        case OP_SIN: un(inline_trans ? "mr_sin(" : "mr_sin_call(", ")"); break;
        case OP_EXP: un(inline_trans ? "mr_exp(" : "mr_exp_call(", ")"); break;
        case OP_LN: un(inline_trans ? "mr_log(" : "mr_log_call(", ")"); break;
        case OP_ADD: bin("", " + ", ""); break;
        case OP_MUL: bin("", " * ", ""); break;
        case OP_MAX: bin("mr_max(", ", ", ")"); break;
        case OP_MIN: bin("mr_min(", ", ", ")"); break;
        case OP_TEX: {
            uint32_t img = n.imm / 4, k = n.imm % 4;
            std::snprintf(buf, sizeof buf, "mr_tex(T[%u].data, T[%u].w, T[%u].h, %uu, ", img, img, img, k);
            bin(buf, ", ", ")");
            break;
        }
        default: s += "0.0"; break;
        }
        s += ";\n";
    }
};

}  // namespace

namespace {
// Values that are exactly +0.0 or 1.0 whatever the pixel: step(..), the constants 0 and 1, products /
// min / max of two such values (AND, AND, OR), and `1 + -(b)` (NOT).  All of these identities are exact
// in IEEE arithmetic (0*1 = +0, 1 + -1 = +0 under round-to-nearest, min/max of {+0, 1} never see a NaN
// or a -0), so evaluating them as boolean logic changes no bit of any channel.  Scenes built from
// Maray's set algebra (range, set_and/or/xor/inv, inside_triangle; reference src/lib.rs:869-913,
// 1094-1097) are mostly made of them: chess has 1 482 steps, 768 min, 256 max.
std::vector<uint8_t> find_booleans(const Program& prog) {
    const size_t n = prog.nodes.size();
    std::vector<uint8_t> kind(n, 0);
    auto is_one = [&](uint32_t v) { return prog.nodes[v].op == OP_CONST && prog.nodes[v].k == 1.0; };
    auto neg_of_bool = [&](uint32_t v) { return prog.nodes[v].op == OP_NEG && kind[prog.nodes[v].a]; };
    for (size_t i = 0; i < n; i++) {
        const Node& nd = prog.nodes[i];
        switch (nd.op) {
        case OP_CONST: {
            uint64_t bits;
            std::memcpy(&bits, &nd.k, 8);
            kind[i] = (bits == 0 || nd.k == 1.0) ? 1 : 0;          // +0.0 only: -0.0 is not a boolean
        } break;
        case OP_STEP: kind[i] = 1; break;
        case OP_MUL: case OP_MIN: case OP_MAX: kind[i] = (kind[nd.a] && kind[nd.b]) ? 1 : 0; break;
        case OP_ADD: kind[i] = ((is_one(nd.a) && neg_of_bool(nd.b)) || (neg_of_bool(nd.a) && is_one(nd.b))) ? 2 : 0; break;
        default: break;
        }
    }
    return kind;
}

// Common generator.  modules[0] always holds the kernel; with opt.separate_segments (and a segmented
// program) every segment function is its own translation unit in modules[1..].
std::vector<std::string> generate(const Program& prog, const CodegenOptions& opt, CodegenInfo* info) {
    // The per-pixel kernel evaluates the values that depend on both x and y; x-only / y-only frontier
    // values are loaded from the column / row tables (each at its first use), the x-only / y-only
    // values behind them are evaluated by the two prologue kernels only.
    const bool hoist = opt.hoist && (!prog.col_values.empty() || !prog.row_values.empty());
    std::vector<uint8_t> load_kind(prog.nodes.size(), 0);
    std::vector<uint32_t> table_index(prog.nodes.size(), 0);
    std::vector<uint32_t> order, order_batch;
    if (hoist) {
        for (size_t k = 0; k < prog.col_values.size(); k++) { load_kind[prog.col_values[k]] = 1; table_index[prog.col_values[k]] = uint32_t(k); }
        for (size_t k = 0; k < prog.row_values.size(); k++) { load_kind[prog.row_values[k]] = 2; table_index[prog.row_values[k]] = uint32_t(k); }
        std::vector<uint8_t> loaded(prog.nodes.size(), 0);
        for (size_t i = 0; i < prog.order.size(); i++) {
            uint32_t id = prog.order[i];
            const Node& n = prog.nodes[id];
            if (n.op == OP_X || n.op == OP_Y) continue;
            if (n.dep == DEP_X || n.dep == DEP_Y) continue;          // prologue work
            auto need = [&](uint32_t v) {
                if (load_kind[v] && !loaded[v]) { loaded[v] = 1; order.push_back(v); order_batch.push_back(0); }
            };
            if (op_is_unary(n.op) || op_is_binary(n.op)) need(n.a);
            if (op_is_binary(n.op)) need(n.b);
            order.push_back(id);
            order_batch.push_back(prog.batch.size() == prog.order.size() ? prog.batch[i] : 0);
        }
        for (int c = 0; c < 3; c++) {
            uint32_t r = prog.root[c];
            if (load_kind[r] && !loaded[r]) { loaded[r] = 1; order.push_back(r); order_batch.push_back(0); }
        }
    } else {
        order = prog.order;
        order_batch = prog.batch.size() == prog.order.size() ? prog.batch : std::vector<uint32_t>(prog.order.size(), 0);
    }
    uint64_t n_trans = prog.stats.op_count[OP_SIN] + prog.stats.op_count[OP_EXP] + prog.stats.op_count[OP_LN];
    // __constant__ table of the scene's constants (64 KiB bank: at most 8000 entries; the rest
    // stay literals).  Stored as bit patterns so NaN/infinity need no special spelling.
    std::vector<int32_t> bank_index(prog.nodes.size(), -1);
    std::vector<uint64_t> bank;
    if (opt.constants_in_bank) {
        for (size_t i = 0; i < prog.nodes.size() && bank.size() < 8000; i++) {
            if (prog.nodes[i].op != OP_CONST) continue;
            // Small integers and halves stay literals: the compiler folds them exactly (x*1, 0+x as
            // a select, compares against 0) and encodes them as 32-bit immediates.
            double kv = prog.nodes[i].k;
            if (kv == kv && std::fabs(kv) <= 256.0 && kv * 2.0 == std::floor(kv * 2.0)) continue;
            uint64_t bits;
            std::memcpy(&bits, &prog.nodes[i].k, 8);
            bank_index[i] = int32_t(bank.size());
            bank.push_back(bits);
        }
    }
    Emitter em{prog, n_trans < opt.inline_transcendentals_below, opt.constants_in_bank ? &bank_index : nullptr};
    const std::vector<uint8_t> booleans = opt.boolean_logic ? find_booleans(prog) : std::vector<uint8_t>();
    if (opt.boolean_logic) em.boolean = &booleans;
    // The NOT form `1 + -(b)` reads b itself, not only its two operands: b must be visible wherever the
    // NOT is evaluated (frame slots and imports of segmented programs).
    auto not_operand = [&](uint32_t id) -> uint32_t {
        if (booleans.empty() || booleans[id] != 2) return UINT32_MAX;
        const Node& n = prog.nodes[id];
        const uint32_t neg = prog.nodes[n.a].op == OP_NEG ? n.a : n.b;
        return prog.nodes[neg].a;
    };
    Emitter em_pre = em;                       // prologue kernels compute hoisted values, never load them
    if (hoist) { em.load_kind = &load_kind; em.table_index = &table_index; }

    const uint32_t seg_len = opt.segment_values ? opt.segment_values : 4096;
    const bool segmented = order.size() > seg_len;
    const uint32_t n_seg = segmented ? uint32_t((order.size() + seg_len - 1) / seg_len) : 1;

    const bool split = opt.separate_segments && segmented;
    const uint32_t block = opt.block ? ((opt.block + 31) / 32) * 32 : 256;
    const bool use_batches = !em.inline_trans;
    const bool scratch = use_batches && opt.scratch_batches;
    em.scratch_batches = scratch;
    std::string prelude = "// generated by maray_b200 (NVRTC back end); compiled with --fmad=false\n";
    // one private instance of the batch helpers per segment function of a single-unit program
    const bool private_helpers = scratch && segmented && !split && opt.private_batch_helpers;
    if (scratch) {
        if (private_helpers) prelude += "#define MR_NO_DEFAULT_BATCH_HELPERS 1\n";
        prelude += "#define MR_SCR_STRIDE " + std::to_string(block) + "\n";
        prelude += "#define MR_BATCH_WIDTH " + std::to_string(opt.batch_width == 4 ? 4 : 2) + "\n";
    }
    prelude += kDeviceSemText;
    prelude += "\n";
    std::string src;
    src.reserve(order.size() * 40 + 8192);
    src += prelude;
    std::string bank_extern;   // what a separately compiled segment needs to see of the constant table
    if (!bank.empty()) {
        src += "__constant__ unsigned long long MRK_BITS[" + std::to_string(bank.size()) + "] = {\n";
        char kb[32];
        for (size_t i = 0; i < bank.size(); i++) {
            std::snprintf(kb, sizeof kb, "0x%016" PRIx64 "ULL,%s", bank[i], (i % 4 == 3) ? "\n" : " ");
            src += kb;
        }
        src += "};\n#define MRK(i) (reinterpret_cast<const double*>(MRK_BITS)[i])\n";
        bank_extern = "extern __constant__ unsigned long long MRK_BITS[" + std::to_string(bank.size()) +
                      "];\n#define MRK(i) (reinterpret_cast<const double*>(MRK_BITS)[i])\n";
    }
    std::vector<std::string> modules;   // modules[0] is filled in at the end
    modules.emplace_back();
    char buf[256];
    uint32_t frame_slots = 0;
    std::vector<int32_t> slot(prog.nodes.size(), -1);

    if (segmented) {
        // Segment of each value, and the last segment that reads it (roots are read by the epilogue).
        std::vector<uint32_t> seg_of(prog.nodes.size(), 0), last_use(prog.nodes.size(), 0);
        for (size_t i = 0; i < order.size(); i++) seg_of[order[i]] = uint32_t(i / seg_len);
        for (size_t i = 0; i < order.size(); i++) {
            const Node& n = prog.nodes[order[i]];
            uint32_t s = seg_of[order[i]];
            auto use = [&](uint32_t id) {
                Op o = prog.nodes[id].op;
                if (o == OP_CONST || o == OP_X || o == OP_Y) return;
                if (last_use[id] < s) last_use[id] = s;
            };
            if (op_is_unary(n.op) || op_is_binary(n.op)) use(n.a);
            if (op_is_binary(n.op)) use(n.b);
            if (uint32_t extra = not_operand(order[i]); extra != UINT32_MAX) use(extra);
        }
        for (int c = 0; c < 3; c++) {
            Op o = prog.nodes[prog.root[c]].op;
            if (o != OP_CONST && o != OP_X && o != OP_Y) last_use[prog.root[c]] = n_seg;   // epilogue
        }
        // Frame slots: a value gets one when it is read after its own segment; slots are recycled
        // once the last reading segment has finished.
        std::vector<std::vector<int32_t>> free_after(n_seg + 2);
        std::vector<int32_t> free_list;
        for (uint32_t s = 0; s < n_seg; s++) {
            if (s > 0) for (int32_t f : free_after[s - 1]) free_list.push_back(f);
            size_t lo = size_t(s) * seg_len, hi = std::min(order.size(), lo + seg_len);
            for (size_t i = lo; i < hi; i++) {
                uint32_t id = order[i];
                Op o = prog.nodes[id].op;
                if (o == OP_X || o == OP_Y) continue;
                if (last_use[id] > s) {
                    int32_t f;
                    if (!free_list.empty()) { f = free_list.back(); free_list.pop_back(); }
                    else f = int32_t(frame_slots++);
                    slot[id] = f;
                    if (last_use[id] <= n_seg) free_after[std::min(last_use[id], n_seg)].push_back(f);
                }
            }
        }

        for (uint32_t s = 0; s < n_seg; s++) {
            size_t lo = size_t(s) * seg_len, hi = std::min(order.size(), lo + seg_len);
            static const char* const kSegArgs =
                "(double* __restrict__ F, const double X, const double Y, const MrTexture* __restrict__ T, "
                "const double* __restrict__ CV, const unsigned int CW, const double* __restrict__ RV, const unsigned int RW)";
            std::string seg_text;
            std::string& out = split ? seg_text : src;
            if (split) {
                // declaration for the kernel's translation unit; the definition gets its own
                std::snprintf(buf, sizeof buf, "extern __device__ void mr_seg%u", s);
                src += buf; src += kSegArgs; src += ";\n";
                seg_text.reserve((hi - lo) * 48 + prelude.size() + 1024);
                seg_text += prelude;
                seg_text += bank_extern;
                std::snprintf(buf, sizeof buf, "__device__ void mr_seg%u", s);
            } else {
                if (private_helpers) {
                    std::snprintf(buf, sizeof buf, "MR_BATCH_HELPERS(_s%u)\n", s);
                    src += buf;
                    std::snprintf(buf, sizeof buf, "_s%u", s);
                    em.helper_suffix = buf;
                }
                std::snprintf(buf, sizeof buf, "__device__ __noinline__ void mr_seg%u", s);
            }
            out += buf; out += kSegArgs; out += " {\n";
            // imports: values defined in earlier segments and read here
            std::vector<uint32_t> imports;
            {
                std::vector<uint32_t> tmp;
                for (size_t i = lo; i < hi; i++) {
                    const Node& n = prog.nodes[order[i]];
                    auto use = [&](uint32_t id) {
                        Op o = prog.nodes[id].op;
                        if (o == OP_CONST || o == OP_X || o == OP_Y) return;
                        if (seg_of[id] < s) tmp.push_back(id);
                    };
                    if (op_is_unary(n.op) || op_is_binary(n.op)) use(n.a);
                    if (op_is_binary(n.op)) use(n.b);
                    if (uint32_t extra = not_operand(order[i]); extra != UINT32_MAX) use(extra);
                }
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                imports.swap(tmp);
            }
            for (uint32_t id : imports) {
                std::snprintf(buf, sizeof buf, "  const double v%u = F[%d];\n", id, slot[id]);
                out += buf;
                em.bool_from_double(out, id);
            }
            for (size_t i = lo; i < hi;) {
                size_t j = i + 1;
                if (use_batches && order_batch[i]) while (j < hi && order_batch[j] == order_batch[i]) j++;
                if (j - i > 1) em.batch_calls(out, order, i, j);
                else em.statement(out, order[i]);
                for (size_t m = i; m < j; m++) {
                    uint32_t id = order[m];
                    if (opt.sync_every && (m - lo) % opt.sync_every == opt.sync_every - 1) out += "  __syncthreads();\n";
                    if (slot[id] >= 0) {
                        std::snprintf(buf, sizeof buf, "  F[%d] = v%u;\n", slot[id], id);
                        out += buf;
                    }
                }
                i = j;
            }
            out += "}\n";
            if (split) modules.push_back(std::move(seg_text));
        }
    }

    if (opt.min_blocks_per_sm)
        std::snprintf(buf, sizeof buf, "extern \"C\" __global__ void __launch_bounds__(%u, %u) %s(const MrParams p) {\n",
                      block, opt.min_blocks_per_sm, kJitKernelName);
    else
        std::snprintf(buf, sizeof buf, "extern \"C\" __global__ void __launch_bounds__(%u) %s(const MrParams p) {\n",
                      block, kJitKernelName);
    src += buf;
    src += "  __shared__ unsigned int stage[3 * " + std::to_string(block) + " / 4];\n";
    src += "  const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;\n";
    src += "  const bool active = j < p.n;\n";
    src += "  const unsigned int pix = p.p0 + (active ? j : 0u);\n";   // idle lanes redo pixel p0
    src += "  const unsigned int yi = pix / p.W;\n";
    src += "  const unsigned int xi = pix - yi * p.W;\n";
    src += "  const double X = (double)xi;\n  const double Y = (double)yi;\n";   // `x as f64`
    src += "  const MrTexture* __restrict__ T = p.tex;\n  (void)T;\n";
    src += "  const double* __restrict__ CV = p.colv + xi;\n  const unsigned int CW = p.W;\n";
    src += "  const double* __restrict__ RV = p.rowv + (yi - p.row_base);\n  const unsigned int RW = p.rows;\n";
    src += "  (void)CV; (void)CW; (void)RV; (void)RW;\n";
    if (segmented) {
        std::snprintf(buf, sizeof buf, "  double F[%u];\n", frame_slots ? frame_slots : 1);
        src += buf;
        for (uint32_t s = 0; s < n_seg; s++) {
            std::snprintf(buf, sizeof buf, "  mr_seg%u(F, X, Y, T, CV, CW, RV, RW);\n", s);
            src += buf;
        }
    } else {
        uint32_t since_sync = 0;
        for (size_t i = 0; i < order.size();) {
            size_t j = i + 1;
            if (use_batches && order_batch[i]) while (j < order.size() && order_batch[j] == order_batch[i]) j++;
            if (j - i > 1) em.batch_calls(src, order, i, j);
            else em.statement(src, order[i]);
            since_sync += uint32_t(j - i);
            if (opt.sync_every && since_sync >= opt.sync_every) { src += "  __syncthreads();\n"; since_sync = 0; }
            i = j;
        }
    }
    src += "  mr_store_block(p, stage, ";
    for (int c = 0; c < 3; c++) {
        uint32_t r = prog.root[c];
        if (segmented && slot[r] >= 0) {
            std::snprintf(buf, sizeof buf, "F[%d]", slot[r]);
            src += buf;
        } else {
            em.operand(src, r);
        }
        src += ", ";
    }
    src += "active, j);\n}\n";

    // Prologue kernels: one thread per column / per row evaluates the x-only / y-only sub-program and
    // stores its frontier values (k-major, so the per-pixel kernel's column loads are coalesced and
    // its row loads are a broadcast).
    if (hoist) {
        for (int which = 1; which <= 2; which++) {
            const bool colk = which == 1;
            src += colk ? "extern \"C\" __global__ void maray_pre_x(double* __restrict__ tab, const unsigned int n, const unsigned int base, const MrTexture* __restrict__ T) {\n"
                        : "extern \"C\" __global__ void maray_pre_y(double* __restrict__ tab, const unsigned int n, const unsigned int base, const MrTexture* __restrict__ T) {\n";
            src += "  const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;\n  if (t >= n) return;\n  (void)T;\n";
            src += colk ? "  const double X = (double)(base + t);\n" : "  const double Y = (double)(base + t);\n";
            // Values that depend on neither coordinate but are not literals either (a texture fetch at
            // constant coordinates and what is computed from it) may feed the x-only / y-only cone: both
            // prologues evaluate them too (there are at most a handful).
            for (uint32_t id : prog.order) {
                const Node& n = prog.nodes[id];
                if (n.op == OP_X || n.op == OP_Y || (n.dep != (colk ? DEP_X : DEP_Y) && n.dep != DEP_CONST)) continue;
                em_pre.statement(src, id);
                if (load_kind[id] == which) {
                    std::snprintf(buf, sizeof buf, "  tab[%uu * n + t] = v%u;\n", table_index[id], id);
                    src += buf;
                }
            }
            src += "}\n";
        }
    }

    if (info) {
        info->segments = n_seg;
        info->frame_slots = segmented ? frame_slots : 0;
        info->transcendentals_inlined = em.inline_trans;
        info->block = block;
        info->n_col = hoist ? uint32_t(prog.col_values.size()) : 0;
        info->n_row = hoist ? uint32_t(prog.row_values.size()) : 0;
        info->dynamic_smem_bytes = scratch ? 2 * kScratchRows * block * uint32_t(sizeof(double)) : 0;   // argument rows + result rows
        info->max_registers = 0;
        if (opt.min_blocks_per_sm) {
            uint32_t r = 65536u / (block * opt.min_blocks_per_sm);
            r &= ~7u;                                   // register allocation granularity
            info->max_registers = r > 255 ? 255 : r;
        }
    }
    modules[0] = std::move(src);
    return modules;
}
}  // namespace

std::string generate_cuda_source(const Program& prog, const CodegenOptions& opt, CodegenInfo* info) {
    CodegenOptions one = opt;
    one.separate_segments = false;
    return generate(prog, one, info)[0];
}

std::vector<std::string> generate_cuda_modules(const Program& prog, const CodegenOptions& opt, CodegenInfo* info) {
    return generate(prog, opt, info);
}

}  // namespace maray
