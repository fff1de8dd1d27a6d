/*
 * sre_lower.cpp -- sre_program_t -> lowered NFA tables (+ DFA when small).
 * See sre_lower.h for the data model and DESIGN.md section 3 for the argument
 * that one step over these tables equals one iteration of the reference's
 * interpreter loop (sre_vm_thompson.c:88-254) as far as the thread *set* goes.
 */
#include "sre_lower.h"

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>

namespace {

inline int byte_kind(unsigned b)
{
    if (b == '\n') return SRE_KIND_NL;
    if (sre_isword(b)) return SRE_KIND_WORD;
    return SRE_KIND_OTHER;
}

const uint8_t ALLOW_ALL = 0x0f;
const uint8_t ALLOW_EOF = 1u << SRE_KIND_EOF;
const uint8_t ALLOW_NL = 1u << SRE_KIND_NL;
const uint8_t ALLOW_WORD = 1u << SRE_KIND_WORD;
const uint8_t ALLOW_NONWORD = ALLOW_ALL & ~ALLOW_WORD;

typedef std::pair<int32_t, uint8_t>  state_key_t;    /* (pc, allow) */

struct closure_builder_t {
    const sre_program_t        *prog;
    std::vector<uint16_t>       visited;    /* per pc: bit `allow` seen      */
    std::vector<int32_t>        touched;
    std::vector<state_key_t>    stack;

    explicit closure_builder_t(const sre_program_t *p)
        : prog(p), visited(p->len + 1, 0) {}

    /*
     * All (pc, allow) reachable from `from` through JMP/SPLIT/SAVE/ASSERT when
     * the previous position is of kind `prev` (SRE_PREV_*).  Mirrors what
     * sre_vm_thompson_add_thread (sre_vm_thompson.c:273-345) plus the
     * assertion_hold continuation (:227-231) add for one position.
     */
    void run(int32_t from, int prev, std::vector<state_key_t> &out)
    {
        out.clear();
        stack.clear();
        stack.push_back(state_key_t(from, ALLOW_ALL));
        bool prev_word = (prev == SRE_PREV_WORD);

        while (!stack.empty()) {
            int32_t pc = stack.back().first;
            uint8_t allow = stack.back().second;
            stack.pop_back();

            for (;;) {
                if ((uint32_t) pc >= prog->len) {
                    break;
                }
                if (visited[pc] & (1u << allow)) {
                    break;
                }
                if (visited[pc] == 0) {
                    touched.push_back(pc);
                }
                visited[pc] |= (uint16_t) (1u << allow);

                const sre_instruction_t &in = prog->insts[pc];
                if (in.opcode == SRE_OPCODE_JMP) {
                    pc = in.x;
                } else if (in.opcode == SRE_OPCODE_SPLIT) {
                    stack.push_back(state_key_t(in.y, allow));
                    pc = in.x;
                } else if (in.opcode == SRE_OPCODE_SAVE) {
                    pc++;
                } else if (in.opcode == SRE_OPCODE_ASSERT) {
                    switch (in.v) {
                    case SRE_REGEX_ASSERT_BIG_A:
                        if (prev != SRE_PREV_START) allow = 0;
                        break;
                    case SRE_REGEX_ASSERT_CARET:
                        if (prev != SRE_PREV_START && prev != SRE_PREV_NL) allow = 0;
                        break;
                    case SRE_REGEX_ASSERT_SMALL_Z:
                        allow &= ALLOW_EOF;
                        break;
                    case SRE_REGEX_ASSERT_DOLLAR:
                        allow &= (ALLOW_EOF | ALLOW_NL);
                        break;
                    case SRE_REGEX_ASSERT_SMALL_B:
                        allow &= prev_word ? ALLOW_NONWORD : ALLOW_WORD;
                        break;
                    case SRE_REGEX_ASSERT_BIG_B:
                        allow &= prev_word ? ALLOW_WORD : ALLOW_NONWORD;
                        break;
                    default:
                        allow = 0;
                        break;
                    }
                    if (allow == 0) {
                        break;
                    }
                    pc++;
                } else {
                    out.push_back(state_key_t(pc, allow));
                    break;
                }
            }
        }
        for (size_t i = 0; i < touched.size(); i++) {
            visited[touched[i]] = 0;
        }
        touched.clear();
    }
};

bool accepts(const sre_program_t *prog, const sre_instruction_t &in, unsigned b)
{
    switch (in.opcode) {
    case SRE_OPCODE_CHAR:
        return in.ch == b;
    case SRE_OPCODE_ANY:
        return true;
    case SRE_OPCODE_IN:
    case SRE_OPCODE_NOTIN: {
        bool hit = false;
        for (uint32_t j = 0; j < in.nranges; j++) {
            const sre_vm_range_t &r = prog->ranges[in.v + j];
            if (b >= r.from && b <= r.to) {
                hit = true;
                break;
            }
        }
        return hit == (in.opcode == SRE_OPCODE_IN);
    }
    default:
        return false;
    }
}

inline void set_bit(uint32_t *w, uint32_t i) { w[i >> 5] |= 1u << (i & 31); }
inline bool get_bit(const uint32_t *w, uint32_t i) { return (w[i >> 5] >> (i & 31)) & 1; }

int build_nfa(const sre_program_t *prog, sre_nfa_t &nfa)
{
    bool lookbehind = false;
    for (uint32_t pc = 0; pc < prog->len; pc++) {
        const sre_instruction_t &in = prog->insts[pc];
        if (in.opcode == SRE_OPCODE_ASSERT
            && (in.v & (SRE_REGEX_ASSERT_BIG_A | SRE_REGEX_ASSERT_CARET
                        | SRE_REGEX_ASSERT_SMALL_B | SRE_REGEX_ASSERT_BIG_B)))
        {
            lookbehind = true;
        }
    }
    nfa.nkinds = lookbehind ? 3 : 1;

    closure_builder_t            cb(prog);
    std::vector<state_key_t>     init_keys, tmp;
    /* closures memoised per (source pc, prev kind) */
    std::map<std::pair<int32_t, int>, std::vector<state_key_t> > memo;
    std::map<state_key_t, uint32_t>  ids;
    std::vector<state_key_t>         work;

    cb.run(0, SRE_PREV_START, init_keys);
    for (size_t i = 0; i < init_keys.size(); i++) {
        if (ids.insert(std::make_pair(init_keys[i], 0u)).second) {
            work.push_back(init_keys[i]);
        }
    }
    while (!work.empty()) {
        state_key_t key = work.back();
        work.pop_back();
        if (prog->insts[key.first].opcode == SRE_OPCODE_MATCH) {
            continue;
        }
        for (uint32_t k = 0; k < nfa.nkinds; k++) {
            std::pair<int32_t, int> mk(key.first, (int) k);
            if (memo.count(mk)) {
                continue;
            }
            cb.run(key.first + 1, nfa.nkinds == 1 ? SRE_PREV_OTHER : (int) k, tmp);
            memo[mk] = tmp;
            for (size_t i = 0; i < tmp.size(); i++) {
                if (ids.insert(std::make_pair(tmp[i], 0u)).second) {
                    work.push_back(tmp[i]);
                }
            }
        }
    }

    /* number the states by (pc, allow): chains pc -> pc+1 become s -> s+1 */
    uint32_t n = 0;
    for (std::map<state_key_t, uint32_t>::iterator it = ids.begin(); it != ids.end(); ++it) {
        it->second = n++;
        nfa.state_pc.push_back(it->first.first);
        nfa.state_allow.push_back(it->first.second);
    }
    nfa.nstates = n;
    nfa.nwords = (n + 31) / 32;
    if (nfa.nwords == 0) {
        nfa.nwords = 1;
    }
    const uint32_t W = nfa.nwords;

    nfa.init.assign(W, 0);
    for (size_t i = 0; i < init_keys.size(); i++) {
        set_bit(&nfa.init[0], ids[init_keys[i]]);
    }

    nfa.follow.assign((size_t) nfa.nkinds * n * W, 0);
    for (uint32_t s = 0; s < n; s++) {
        if (prog->insts[nfa.state_pc[s]].opcode == SRE_OPCODE_MATCH) {
            continue;
        }
        for (uint32_t k = 0; k < nfa.nkinds; k++) {
            const std::vector<state_key_t> &c = memo[std::make_pair(nfa.state_pc[s], (int) k)];
            uint32_t *row = &nfa.follow[((size_t) k * n + s) * W];
            for (size_t i = 0; i < c.size(); i++) {
                set_bit(row, ids[c[i]]);
            }
        }
    }

    /* per byte: movers / surviving matches; then group bytes into classes */
    std::vector<uint32_t> mv256((size_t) 256 * W, 0), mt256((size_t) 256 * W, 0);
    nfa.mt_eof.assign(W, 0);
    for (uint32_t s = 0; s < n; s++) {
        const sre_instruction_t &in = prog->insts[nfa.state_pc[s]];
        uint8_t allow = nfa.state_allow[s];
        if (in.opcode == SRE_OPCODE_MATCH) {
            if (allow & ALLOW_EOF) {
                set_bit(&nfa.mt_eof[0], s);
            }
            if (allow != ALLOW_ALL) {
                nfa.has_match_lookahead = true;
            }
        }
        for (unsigned b = 0; b < 256; b++) {
            if (!(allow & (1u << byte_kind(b)))) {
                continue;
            }
            if (in.opcode == SRE_OPCODE_MATCH) {
                set_bit(&mt256[(size_t) b * W], s);
            } else if (accepts(prog, in, b)) {
                set_bit(&mv256[(size_t) b * W], s);
            }
        }
    }

    std::map<std::string, uint32_t> sigs;
    for (unsigned b = 0; b < 256; b++) {
        std::string sig((const char *) &mv256[(size_t) b * W], W * 4);
        sig.append((const char *) &mt256[(size_t) b * W], W * 4);
        sig.push_back(nfa.nkinds == 1 ? 0 : (char) (byte_kind(b) - 1));
        std::map<std::string, uint32_t>::iterator it = sigs.find(sig);
        if (it == sigs.end()) {
            uint32_t c = (uint32_t) sigs.size();
            sigs[sig] = c;
            nfa.clsmap[b] = (uint8_t) c;
            nfa.cls_kind.push_back(nfa.nkinds == 1 ? 0 : (uint8_t) (byte_kind(b) - 1));
            nfa.mv.insert(nfa.mv.end(), &mv256[(size_t) b * W], &mv256[(size_t) b * W] + W);
            nfa.mt.insert(nfa.mt.end(), &mt256[(size_t) b * W], &mt256[(size_t) b * W] + W);
        } else {
            nfa.clsmap[b] = (uint8_t) it->second;
        }
    }
    nfa.nclasses = (uint32_t) sigs.size();

    nfa.shift_mask.assign(W, 0);
    for (uint32_t s = 0; s + 1 < n; s++) {
        if (prog->insts[nfa.state_pc[s]].opcode == SRE_OPCODE_MATCH) {
            continue;
        }
        bool only_next = true;
        for (uint32_t k = 0; k < nfa.nkinds && only_next; k++) {
            const uint32_t *row = &nfa.follow[((size_t) k * n + s) * W];
            for (uint32_t w = 0; w < W; w++) {
                uint32_t want = ((s + 1) >> 5) == w ? 1u << ((s + 1) & 31) : 0;
                if (row[w] != want) {
                    only_next = false;
                    break;
                }
            }
        }
        if (only_next) {
            set_bit(&nfa.shift_mask[0], s);
        }
    }
    return SRE_OK;
}

struct set_hash_t {
    size_t operator()(const std::vector<uint32_t> &v) const {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < v.size(); i++) {
            h = (h ^ v[i]) * 1099511628211ull;
        }
        return (size_t) h;
    }
};

/* subset construction with the exact step semantics of sre_lower.h */
bool build_dfa(const sre_nfa_t &nfa, uint32_t max_states, sre_dfa_t &dfa)
{
    const uint32_t W = nfa.nwords, C = nfa.nclasses, n = nfa.nstates;
    std::unordered_map<std::vector<uint32_t>, uint32_t, set_hash_t> ids;
    std::vector<std::vector<uint32_t> > sets;
    std::vector<uint16_t> trans;        /* [state][C] over NFA classes */
    std::vector<uint8_t> restart;       /* [state][C]: only the `any` thread moved */
    const uint32_t ACC = 1;
    /* the state of the ".*?" prefix: pc 1, no pending look-ahead */
    int32_t any_state = -1;
    for (uint32_t s = 0; s < n; s++) {
        if (nfa.state_pc[s] == 1 && nfa.state_allow[s] == 0x0f) {
            any_state = (int32_t) s;
        }
    }

    sets.push_back(nfa.init);           /* 0 = start */
    ids[nfa.init] = 0;
    sets.push_back(std::vector<uint32_t>());    /* 1 = ACC (no set) */
    trans.assign(2 * C, (uint16_t) ACC);
    restart.assign(2 * C, 0);

    std::vector<uint32_t> next(W);
    for (uint32_t cur = 0; cur < sets.size(); cur++) {
        if (cur == ACC) {
            continue;
        }
        for (uint32_t c = 0; c < C; c++) {
            const std::vector<uint32_t> S = sets[cur];
            const uint32_t *mv = &nfa.mv[(size_t) c * W], *mt = &nfa.mt[(size_t) c * W];
            bool matched = false;
            for (uint32_t w = 0; w < W; w++) {
                if (S[w] & mt[w]) {
                    matched = true;
                    break;
                }
            }
            if (matched) {
                trans[(size_t) cur * C + c] = (uint16_t) ACC;
                continue;
            }
            std::fill(next.begin(), next.end(), 0u);
            bool only_any = any_state >= 0;
            for (uint32_t w = 0; w < W; w++) {
                uint32_t m = S[w] & mv[w];
                const uint32_t any_bit = (any_state >= 0 && (uint32_t) any_state / 32 == w)
                                             ? 1u << (any_state & 31) : 0;
                if (m != any_bit) {
                    only_any = false;
                }
                while (m) {
                    uint32_t s = w * 32 + __builtin_ctz(m);
                    m &= m - 1;
                    if (s >= n) {
                        break;
                    }
                    const uint32_t *row = nfa.follow_row(nfa.cls_kind[c], s);
                    for (uint32_t x = 0; x < W; x++) {
                        next[x] |= row[x];
                    }
                }
            }
            std::unordered_map<std::vector<uint32_t>, uint32_t, set_hash_t>::iterator it = ids.find(next);
            uint32_t id;
            if (it == ids.end()) {
                id = (uint32_t) sets.size();
                if (id >= max_states || id >= 65535) {
                    return false;
                }
                ids[next] = id;
                sets.push_back(next);
                trans.resize(trans.size() + C, 0);
                restart.resize(restart.size() + C, 0);
            } else {
                id = it->second;
            }
            trans[(size_t) cur * C + c] = (uint16_t) id;
            restart[(size_t) cur * C + c] = only_any ? 1 : 0;
        }
    }

    /* the EOF step (mt_eof) per subset state */
    const uint32_t D0 = (uint32_t) sets.size();
    std::vector<uint8_t> fin0(D0, 0);
    fin0[ACC] = 1;
    for (uint32_t d = 0; d < D0; d++) {
        if (d == ACC) {
            continue;
        }
        for (uint32_t w = 0; w < W; w++) {
            if (sets[d][w] & nfa.mt_eof[w]) {
                fin0[d] = 1;
                break;
            }
        }
    }

    /*
     * Minimise (Moore refinement).  Two states are merged when no input tells
     * them apart by: being ACC (kept on its own: kernels stop / report at the
     * step that enters it), the EOF verdict, and per byte class the successor
     * block and the restart flag of the transition (the Pike start hint reads
     * it).  Block of state 0 stays 0, ACC stays 1.
     */
    std::vector<uint32_t> block(D0);
    uint32_t nblocks = 0;
    {
        bool third = false;
        for (uint32_t d = 0; d < D0; d++) {
            block[d] = d == ACC ? 1 : (fin0[d] == fin0[0] ? 0 : 2);
            third |= block[d] == 2;
        }
        /* (the number of blocks there really are: the refinement stops when a round adds none,
         * and must not take a first round that ends with three for a stable one) */
        nblocks = third ? 3 : 2;
        /* state 0 and ACC are visited first: their blocks keep the numbers 0 and 1 */
        std::vector<uint32_t> order;
        order.push_back(0);
        order.push_back(ACC);
        for (uint32_t d = 0; d < D0; d++) {
            if (d != 0 && d != ACC) {
                order.push_back(d);
            }
        }
        /* signature of d under the current blocks: (block, per class successor block + flag);
         * states are bucketed by a hash of it and compared in full within a bucket */
        auto sig_at = [&](uint32_t d, uint32_t c) -> uint32_t {
            return c == 0 ? block[d]
                          : block[trans[(size_t) d * C + c - 1]] * 2 + restart[(size_t) d * C + c - 1];
        };
        std::vector<uint32_t> newblock(D0), rep;
        std::unordered_map<uint64_t, std::vector<uint32_t> > buckets;
        for (;;) {
            buckets.clear();
            rep.clear();
            for (uint32_t d : order) {
                uint64_t h = 1469598103934665603ull;
                for (uint32_t c = 0; c <= C; c++) {
                    h = (h ^ sig_at(d, c)) * 1099511628211ull;
                }
                std::vector<uint32_t> &cands = buckets[h];
                uint32_t found = 0xffffffffu;
                for (uint32_t b : cands) {
                    bool same = true;
                    for (uint32_t c = 0; c <= C && same; c++) {
                        same = sig_at(rep[b], c) == sig_at(d, c);
                    }
                    if (same) {
                        found = b;
                        break;
                    }
                }
                if (found == 0xffffffffu) {
                    found = (uint32_t) rep.size();
                    rep.push_back(d);
                    cands.push_back(found);
                }
                newblock[d] = found;
            }
            const uint32_t nb = (uint32_t) rep.size();
            const bool stable = (nb == nblocks);
            block = newblock;
            nblocks = nb;
            if (stable) {
                break;
            }
        }
    }
    if (nblocks < D0) {
        std::vector<uint16_t> t2((size_t) nblocks * C, 0);
        std::vector<uint8_t> r2((size_t) nblocks * C, 0), f2(nblocks, 0);
        for (uint32_t d = 0; d < D0; d++) {
            const uint32_t bd = block[d];
            f2[bd] = fin0[d];
            for (uint32_t c = 0; c < C; c++) {
                t2[(size_t) bd * C + c] = (uint16_t) block[trans[(size_t) d * C + c]];
                r2[(size_t) bd * C + c] = restart[(size_t) d * C + c];
            }
        }
        trans.swap(t2);
        restart.swap(r2);
        fin0.swap(f2);
    }

    const uint32_t D = nblocks < D0 ? nblocks : D0;
    dfa.nstates = D;
    dfa.start = 0;
    dfa.acc = ACC;
    dfa.fin.assign(fin0.begin(), fin0.begin() + D);

    if (D <= 128) {
        dfa.h256.assign((size_t) 256 * 256, 0);
        for (uint32_t d = 0; d < D; d++) {
            for (unsigned b = 0; b < 256; b++) {
                const uint32_t c = nfa.clsmap[b];
                const uint8_t e = (uint8_t) (trans[(size_t) d * C + c]
                                             | (restart[(size_t) d * C + c] ? 0x80 : 0));
                dfa.h256[(size_t) d * 256 + b] = e;
                dfa.h256[(size_t) (d + 128) * 256 + b] = e;
            }
        }
    }

    if (D <= 32767) {
        dfa.hncls = C;
        for (unsigned b = 0; b < 256; b++) {
            dfa.hclsmap[b] = nfa.clsmap[b];
        }
        dfa.hcls.assign((size_t) D * C, 0);
        for (uint32_t d = 0; d < D; d++) {
            for (uint32_t c = 0; c < C; c++) {
                dfa.hcls[(size_t) d * C + c] =
                    (uint16_t) (trans[(size_t) d * C + c] | (restart[(size_t) d * C + c] ? 0x8000 : 0));
            }
        }
    }

    /* merge NFA byte classes whose DFA columns coincide */
    std::map<std::vector<uint16_t>, uint32_t> cols;
    std::vector<uint32_t> remap(C);
    std::vector<uint32_t> rep;
    for (uint32_t c = 0; c < C; c++) {
        std::vector<uint16_t> col(D);
        for (uint32_t d = 0; d < D; d++) {
            col[d] = trans[(size_t) d * C + c];
        }
        std::map<std::vector<uint16_t>, uint32_t>::iterator it = cols.find(col);
        if (it == cols.end()) {
            remap[c] = (uint32_t) cols.size();
            cols[col] = remap[c];
            rep.push_back(c);
        } else {
            remap[c] = it->second;
        }
    }
    dfa.nclasses = (uint32_t) rep.size();
    for (unsigned b = 0; b < 256; b++) {
        dfa.clsmap[b] = (uint8_t) remap[nfa.clsmap[b]];
    }
    dfa.trans.assign((size_t) D * dfa.nclasses, 0);
    for (uint32_t d = 0; d < D; d++) {
        for (uint32_t c = 0; c < dfa.nclasses; c++) {
            dfa.trans[(size_t) d * dfa.nclasses + c] = trans[(size_t) d * C + rep[c]];
        }
    }
    if (D <= 256) {
        dfa.t256.assign((size_t) D * 256, 0);
        for (uint32_t d = 0; d < D; d++) {
            for (unsigned b = 0; b < 256; b++) {
                dfa.t256[(size_t) d * 256 + b] =
                    (uint8_t) dfa.trans[(size_t) d * dfa.nclasses + dfa.clsmap[b]];
            }
        }
    }
    return true;
}

}  // namespace

int sre_lower_program(const sre_program_t *prog, uint32_t max_dfa_states,
    sre_lowered_t *out)
{
    if (prog == NULL || prog->magic != SRE_PROGRAM_MAGIC || prog->len == 0) {
        return SRE_ERROR;
    }
    if (build_nfa(prog, out->nfa) != SRE_OK) {
        return SRE_ERROR;
    }
    out->has_dfa = false;
    if (max_dfa_states >= 2) {
        out->has_dfa = build_dfa(out->nfa, max_dfa_states, out->dfa);
        if (!out->has_dfa) {
            out->dfa = sre_dfa_t();
        }
    }
    return SRE_OK;
}
