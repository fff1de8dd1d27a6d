/*
 * sre_pdfa.cpp -- subset construction over ordered Pike thread lists (see
 * sre_pdfa.h).  The step below is the step of the closure-table Pike
 * (kernels/sre_pike_table.cu, host model oracle/lower_check.cpp:
 * lc_table_pike) without its next-byte pruning: a pruned thread is one that
 * dies on the next byte anyway, and an instruction that prunes one thread
 * prunes every thread parked on it, so the pruning changes no result.
 */
#include "sre_pdfa.h"

#include <functional>
#include <map>
#include <stdlib.h>
#include <string.h>

namespace {

inline bool isword(uint32_t c)
{
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}

/* one thread of a list: the instruction it is parked on | 0x8000 when it is parked on \b / \B and
 * the byte in front of it was a word byte (sre_vm_pike.c:868-887: t->seen_word) */
typedef std::vector<uint16_t> list_t;

struct item_t {             /* a thread being stepped */
    uint16_t ent;           /* park | seen_word << 15 */
    uint8_t  parent;        /* index in the list it descends from */
    uint32_t mask0;         /* slots SAVEd at the position of the step by resolved look-aheads */
};

}  // namespace

bool sre_build_pdfa(const sre_program_t *prog, const sre_closure_table_t &T, uint32_t max_states, sre_pdfa_t &out)
{
    const uint32_t np = T.npark;
    if (np == 0 || np >= 0x8000 || max_states > 0x7fff) {
        return false;
    }
    bool lookahead = false, wordy = false;
    for (uint32_t P = 0; P < np; P++) {
        lookahead |= T.kind[P] >= 2;
        wordy |= T.kind[P] >= 4;
    }
    out.ctx_dep = T.ctx_dep;
    out.lookahead = lookahead;
    out.max_slots = T.max_slots;
    /*
     * SRE_PDFA_EXACT=1 (programs with look-ahead assertions; off by default: validated on the CPU
     * model only, see DESIGN.md 3.3 "known divergence"): the step below walks the bytecode itself
     * with the reference's one tag word per instruction, instead of combining the closure tables,
     * and the instructions that carry the list's tag are part of the state.  Exact also where one
     * closure parks a look-ahead assertion and visits what lies behind it (`(\B)?x`).
     */
    const char *exact_env = getenv("SRE_PDFA_EXACT");
    const bool exact = lookahead && exact_env != nullptr && atoi(exact_env) != 0;
    std::vector<int32_t> pc_park(prog->len, -1);
    for (uint32_t P = 0; P < np; P++) {
        pc_park[T.park_pc[P]] = (int32_t) P;
    }
    std::vector<uint32_t> slot_base(prog->len, 0);          /* first slot of the regex that owns a pc */
    {
        uint32_t r = 0, base = 0;
        for (uint32_t pc = 0; pc < prog->len; pc++) {
            slot_base[pc] = base;
            if (prog->insts[pc].opcode == SRE_OPCODE_MATCH && r + 1 < prog->nregexes) {
                base += 2 * (uint32_t) (prog->multi_ncaps[r] + 1);
                r++;
            }
        }
    }
    /* tag words of the walk: 0 = older, 1 = the step the current list was built in, 2 = this step */
    std::vector<uint8_t> tagw(prog->len, 0);
    struct walked_t {
        uint16_t ent;           /* parked number | seen_word << 15 */
        uint32_t mask;          /* slots SAVEd on the way */
    };
    /* add_thread (sre_vm_pike.c:756-942) under tag word `tag`; prev: 0 offset 0, 1 after a newline,
     * 2 after a word byte, 3 after anything else.  -> true when MATCH was reached with want_done
     * (its slots in *done_mask, its instruction in *done_pc) */
    std::function<bool(int32_t, uint32_t, uint8_t, uint32_t, bool, std::vector<walked_t> &, uint32_t *, int32_t *)> walk =
        [&](int32_t pc, uint32_t mask, uint8_t tag, uint32_t prev, bool want_done, std::vector<walked_t> &outl,
            uint32_t *done_mask, int32_t *done_pc) -> bool {
        const sre_instruction_t &in = prog->insts[pc];
        if (tagw[pc] == tag) {
            if (in.opcode == SRE_OPCODE_SPLIT && tagw[in.y] != tag) {       /* the revisited-SPLIT rule */
                return walk(in.y, mask, tag, prev, want_done, outl, done_mask, done_pc);
            }
            return false;
        }
        tagw[pc] = tag;
        bool sw = false;
        switch (in.opcode) {
        case SRE_OPCODE_JMP:
            return walk(in.x, mask, tag, prev, want_done, outl, done_mask, done_pc);
        case SRE_OPCODE_SPLIT:
            if (walk(in.x, mask, tag, prev, want_done, outl, done_mask, done_pc)) {
                return true;
            }
            return walk(in.y, mask, tag, prev, want_done, outl, done_mask, done_pc);
        case SRE_OPCODE_SAVE:
            return walk(pc + 1, mask | (1u << ((uint32_t) in.v - slot_base[pc])), tag, prev, want_done, outl,
                        done_mask, done_pc);
        case SRE_OPCODE_ASSERT:
            if (in.v == SRE_REGEX_ASSERT_BIG_A) {
                return prev == 0 && T.ctx_dep ? walk(pc + 1, mask, tag, prev, want_done, outl, done_mask, done_pc)
                                              : false;
            }
            if (in.v == SRE_REGEX_ASSERT_CARET) {
                return (prev == 0 || prev == 1) ? walk(pc + 1, mask, tag, prev, want_done, outl, done_mask, done_pc)
                                                : false;
            }
            sw = (in.v == SRE_REGEX_ASSERT_SMALL_B || in.v == SRE_REGEX_ASSERT_BIG_B) && prev == 2;
            break;
        case SRE_OPCODE_MATCH:
            if (want_done) {
                *done_mask = mask;
                *done_pc = pc;
                return true;
            }
            break;
        default:
            break;
        }
        walked_t w;
        w.ent = (uint16_t) ((uint32_t) pc_park[pc] | (sw ? 0x8000u : 0u));
        w.mask = mask;
        outl.push_back(w);
        return false;
    };
    /* what a state must remember of the byte in front of it: 0 nothing / offset 0, 1 a newline,
     * 2 a word byte, 3 anything else -- only as far as some closure can tell the difference */
    auto kind_of = [&](uint32_t b) -> uint32_t {
        if (wordy && isword(b)) {
            return 2;
        }
        if (T.ctx_dep) {
            return b == '\n' ? 1u : 3u;
        }
        return 0;               /* nothing to tell apart from "nothing in front" */
    };
    auto ctx_of = [&](uint32_t pk) -> uint32_t { return !T.ctx_dep ? 0u : pk == 0 ? 0u : pk == 1 ? 1u : 2u; };

    /* byte classes: bytes that the same parked instructions accept and that leave the same
     * look-behind kind (and, with `$`, '\n' apart) */
    {
        std::map<std::vector<uint8_t>, uint32_t> sigs;
        for (uint32_t b = 0; b < 256; b++) {
            std::vector<uint8_t> sig(T.nsets + 2);
            for (uint32_t k = 0; k < T.nsets; k++) {
                sig[k] = (T.accept[(size_t) k * 8 + (b >> 5)] >> (b & 31)) & 1;
            }
            sig[T.nsets] = (uint8_t) kind_of(b);
            sig[T.nsets + 1] = lookahead && b == '\n';
            std::map<std::vector<uint8_t>, uint32_t>::iterator it = sigs.find(sig);
            if (it == sigs.end()) {
                it = sigs.insert(std::make_pair(sig, (uint32_t) sigs.size())).first;
            }
            out.clsmap[b] = (uint8_t) it->second;
        }
        out.nclasses = (uint32_t) sigs.size();
    }
    const uint32_t C = out.nclasses;
    std::vector<uint32_t> rep(C, 0);
    for (int b = 255; b >= 0; b--) {
        rep[out.clsmap[b]] = (uint32_t) b;
    }
    auto accepts = [&](uint32_t P, uint32_t b) -> bool {
        return (T.accept[(size_t) T.acc_idx[P] * 8 + (b >> 5)] >> (b & 31)) & 1;
    };

    /*
     * A state = (look-behind kind, list, odd marks).  The reference dedups closures by one tag
     * word per instruction, and a look-ahead thread that holds appends its closure under the tag
     * of the PREVIOUS step (sre_vm_pike.c:484-509: tag--), so what that step tagged is part of
     * the state.  Mostly that is "the instructions of the list" -- but not quite: a MATCH
     * reached by a closure is tagged and not parked (:889-899), and an instruction appended to
     * the next list and then visited again by a held closure carries the older tag from then on.
     * `odd` = the parked instructions whose mark differs from their membership in the list
     * (programs with look-ahead assertions only; nothing else ever reads the older marks).
     */
    struct key_t {
        uint32_t first;
        list_t   second, odd;
        key_t() : first(0) {}
        key_t(uint32_t k, const list_t &l, const list_t &o) : first(k), second(l), odd(o) {}
        bool operator<(const key_t &r) const
        {
            return first != r.first ? first < r.first : second != r.second ? second < r.second : odd < r.odd;
        }
    };
    std::map<key_t, uint32_t> ids;
    std::vector<key_t> states;
    states.push_back(key_t(0, list_t(), list_t()));     /* state 0: the empty list */
    ids[states[0]] = 0;
    auto intern = [&](uint32_t pk, const list_t &l, const list_t &odd, uint32_t *id) -> bool {
        /* the empty list is one state whatever came before it */
        const key_t k(l.empty() ? 0u : pk, l, l.empty() ? list_t() : odd);
        std::map<key_t, uint32_t>::iterator it = ids.find(k);
        if (it != ids.end()) {
            *id = it->second;
            return true;
        }
        if (states.size() >= max_states) {
            return false;
        }
        *id = (uint32_t) states.size();
        ids[k] = *id;
        states.push_back(k);
        return true;
    };

    /* the start closures (closure of "pc 0", number np) by what lies in front of the first byte */
    for (uint32_t pk = 0; pk < 4; pk++) {
        list_t init;
        std::vector<uint8_t> marks(np, 0);
        out.init_mask_ofs[pk] = (uint32_t) out.init_mask.size();
        const uint32_t vofs = ctx_of(pk) * (np + 2);
        list_t tagged;          /* exact: the instructions the walk tagged */
        if (exact) {
            std::vector<walked_t> w;
            uint32_t dm = 0;
            int32_t dp = 0;
            std::fill(tagw.begin(), tagw.end(), (uint8_t) 0);
            /* (a context the program cannot tell from "elsewhere" is walked as the kind it collapses to) */
            const uint32_t eff0 = pk == 0 ? 0u : kind_of(pk == 1 ? '\n' : pk == 2 ? 'a' : '.');
            walk(0, 0, 1, pk == 0 ? 0u : (eff0 == 0 ? 3u : eff0), false, w, &dm, &dp);
            for (size_t k = 0; k < w.size(); k++) {
                init.push_back(w[k].ent);
                out.init_mask.push_back(w[k].mask);
            }
            for (uint32_t pc = 0; pc < prog->len; pc++) {
                if (tagw[pc] == 1) {
                    tagged.push_back((uint16_t) pc);
                }
            }
        } else
        for (uint32_t e = T.ofs[vofs + np]; e < T.ofs[vofs + np + 1]; e++) {
            const uint32_t fp = T.ent[e];
            if (marks[fp]) {
                continue;
            }
            marks[fp] = 1;
            init.push_back((uint16_t) (fp | ((T.kind[fp] >= 4 && pk == 2) ? 0x8000u : 0u)));
            out.init_mask.push_back(T.emask[e]);
        }
        if (init.empty() || init.size() > 255) {
            return false;
        }
        /* (kinds the program cannot tell apart give the same state) */
        const uint32_t eff = pk == 0 ? 0u : kind_of(pk == 1 ? '\n' : pk == 2 ? 'a' : '.');
        if (!intern(pk == 0 ? 0u : eff, init, tagged, &out.init[pk])) {
            return false;
        }
    }

    std::vector<uint8_t> m_prev(np), m_cur(np);
    /*
     * One step of state s: on byte b (eof = false) or at the end of the input.  Threads are taken
     * in list order; a look-ahead thread whose assertion holds is replaced, in place, by its
     * closure at the same position (sre_vm_pike.c:505-523, dedup against the previous step's
     * marks: the tag-- trick); a consuming thread that takes b appends its closure at the next
     * position; MATCH (parked, or reached by a closure that may report it) ends the step and
     * cuts what has lower priority.
     */
    struct step_out_t {
        list_t                next, odd;
        std::vector<uint8_t>  parents;
        std::vector<uint32_t> mask0, mask1;
        bool                  mev = false;
        uint8_t               mpar = 0;
        uint32_t              mmask0 = 0, mmask1 = 0;
        uint16_t              mreg = 0;
    };
    auto step = [&](const key_t &st, bool eof, uint32_t b, step_out_t &o) {
        const list_t &cur = st.second;
        const uint32_t pk = st.first;
        if (exact) {
            /* the reference's step on the bytecode: st.odd = the instructions tagged by the step
             * that built the list */
            std::fill(tagw.begin(), tagw.end(), (uint8_t) 0);
            for (size_t j = 0; j < st.odd.size(); j++) {
                tagw[st.odd[j]] = 1;
            }
            const bool word_b = !eof && isword(b);
            const uint32_t next_prev = b == '\n' ? 1u : word_b ? 2u : 3u;
            /* (the kind of the byte in front, as far as the program can tell: 0 stands for
             * "elsewhere" too when nothing looks at offset 0 or at newlines) */
            const uint32_t hold_prev = (pk == 0 && !T.ctx_dep) ? 3u : pk;
            std::vector<item_t> holdx;
            std::vector<walked_t> w;
            size_t ix = 0;
            auto done = [&]() {
                for (uint32_t pc = 0; pc < prog->len; pc++) {
                    if (tagw[pc] == 2) {
                        o.odd.push_back((uint16_t) pc);
                    }
                }
            };
            for (;;) {
                item_t t;
                if (!holdx.empty()) {
                    t = holdx.back();
                    holdx.pop_back();
                } else if (ix < cur.size()) {
                    t.ent = cur[ix];
                    t.parent = (uint8_t) ix;
                    t.mask0 = 0;
                    ix++;
                } else {
                    break;
                }
                const uint32_t P = t.ent & 0x7fff;
                const bool sw = (t.ent & 0x8000) != 0;
                const uint32_t kind = T.kind[P];
                const int32_t pc = T.park_pc[P];
                uint32_t dm = 0;
                int32_t dp = 0;
                if (kind >= 2) {
                    bool holds;
                    switch (kind) {
                    case 2:  holds = eof; break;
                    case 3:  holds = eof || b == '\n'; break;
                    case 4:  holds = (sw == word_b); break;
                    default: holds = (sw != word_b); break;
                    }
                    if (!holds) {
                        continue;
                    }
                    w.clear();
                    walk(pc + 1, t.mask0, 1, hold_prev, false, w, &dm, &dp);
                    for (size_t k = w.size(); k-- > 0;) {       /* prepended: the first one on top */
                        item_t a;
                        a.ent = w[k].ent;
                        a.parent = t.parent;
                        a.mask0 = w[k].mask;
                        holdx.push_back(a);
                    }
                } else if (kind == 1) {
                    o.mev = true;
                    o.mpar = t.parent;
                    o.mmask0 = t.mask0;
                    o.mmask1 = 0;
                    o.mreg = T.regex[P];
                    done();
                    return;
                } else if (!eof && accepts(P, b)) {
                    w.clear();
                    const bool hit = walk(pc + 1, 0, 2, next_prev, true, w, &dm, &dp);
                    for (size_t k = 0; k < w.size(); k++) {
                        o.next.push_back(w[k].ent);
                        o.parents.push_back(t.parent);
                        o.mask0.push_back(t.mask0);
                        o.mask1.push_back(w[k].mask);
                    }
                    if (hit) {
                        o.mev = true;
                        o.mpar = t.parent;
                        o.mmask0 = t.mask0;
                        o.mmask1 = dm;
                        o.mreg = T.regex[pc_park[dp]];
                        done();
                        return;
                    }
                }
            }
            done();
            return;
        }
        const bool prev_word = pk == 2, cur_word = !eof && isword(b);
        const uint32_t vhold = ctx_of(pk) * (np + 2);
        const uint32_t vnext = (T.ctx_dep ? (b == '\n' ? 1u : 2u) : 0u) * (np + 2);
        memset(m_cur.data(), 0, np);
        memset(m_prev.data(), 0, np);
        for (size_t j = 0; j < cur.size(); j++) {
            m_prev[cur[j] & 0x7fff] = 1;
        }
        for (size_t j = 0; j < st.odd.size(); j++) {
            m_prev[st.odd[j]] ^= 1;
        }
        /* what this step leaves tagged, against the list it leaves */
        auto finish = [&]() {
            if (!lookahead) {
                return;
            }
            std::vector<uint8_t> in_list(np, 0);
            for (size_t j = 0; j < o.next.size(); j++) {
                in_list[o.next[j] & 0x7fff] = 1;
            }
            for (uint32_t P = 0; P < np; P++) {
                if (m_cur[P] != in_list[P]) {
                    o.odd.push_back((uint16_t) P);
                }
            }
        };
        std::vector<item_t> hold;           /* LIFO, back = top */
        size_t i = 0;
        for (;;) {
            item_t t;
            if (!hold.empty()) {
                t = hold.back();
                hold.pop_back();
            } else if (i < cur.size()) {
                t.ent = cur[i];
                t.parent = (uint8_t) i;
                t.mask0 = 0;
                i++;
            } else {
                break;
            }
            const uint32_t P = t.ent & 0x7fff;
            const bool sw = (t.ent & 0x8000) != 0;
            const uint32_t kind = T.kind[P];
            if (kind >= 2) {
                bool holds;
                switch (kind) {
                case 2:  holds = eof; break;
                case 3:  holds = eof || b == '\n'; break;
                case 4:  holds = (sw == cur_word); break;
                default: holds = (sw != cur_word); break;
                }
                if (!holds) {
                    continue;
                }
                std::vector<item_t> add;
                for (uint32_t e = T.ofs[vhold + P]; e < T.ofs[vhold + P + 1]; e++) {
                    const uint32_t fp = T.ent[e];
                    if (m_prev[fp]) {
                        continue;
                    }
                    m_prev[fp] = 1;
                    m_cur[fp] = 0;
                    item_t a;
                    a.ent = (uint16_t) (fp | ((T.kind[fp] >= 4 && prev_word) ? 0x8000u : 0u));
                    a.parent = t.parent;
                    a.mask0 = t.mask0 | T.emask[e];
                    add.push_back(a);
                }
                for (size_t k = add.size(); k-- > 0;) {     /* prepended: the first one on top */
                    hold.push_back(add[k]);
                }
            } else if (kind == 1) {
                /* a parked MATCH thread: sre_vm_pike.c:535-553 */
                o.mev = true;
                o.mpar = t.parent;
                o.mmask0 = t.mask0;
                o.mmask1 = 0;
                o.mreg = T.regex[P];
                finish();
                return;
            } else if (!eof && accepts(P, b)) {
                for (uint32_t e = T.ofs[vnext + P]; e < T.ofs[vnext + P + 1]; e++) {
                    const uint32_t fp = T.ent[e];
                    if (m_cur[fp]) {
                        continue;
                    }
                    m_cur[fp] = 1;
                    m_prev[fp] = 0;
                    if (T.kind[fp] == 1) {
                        /* the closure reached MATCH: report and cut what has lower priority */
                        o.mev = true;
                        o.mpar = t.parent;
                        o.mmask0 = t.mask0;
                        o.mmask1 = T.emask[e];
                        o.mreg = T.regex[fp];
                        finish();
                        return;
                    }
                    o.next.push_back((uint16_t) (fp | ((T.kind[fp] >= 4 && cur_word) ? 0x8000u : 0u)));
                    o.parents.push_back(t.parent);
                    o.mask0.push_back(t.mask0);
                    o.mask1.push_back(T.emask[e]);
                }
            }
        }
        finish();
    };

    for (uint32_t s = 0; s < states.size(); s++) {
        for (uint32_t c = 0; c < C; c++) {
            const key_t cur = states[s];            /* copy: states grows below */
            const uint32_t b = rep[c];
            step_out_t o;
            step(cur, false, b, o);
            if (o.next.size() > 255) {
                return false;
            }
            uint32_t id;
            if (!intern(kind_of(b), o.next, o.odd, &id)) {
                return false;
            }
            out.trans.push_back((uint16_t) (id | (o.mev ? 0x8000u : 0u)));
            out.eofs.push_back((uint32_t) out.eparent.size());
            out.eparent.insert(out.eparent.end(), o.parents.begin(), o.parents.end());
            out.emask.insert(out.emask.end(), o.mask1.begin(), o.mask1.end());
            out.emask0.insert(out.emask0.end(), o.mask0.begin(), o.mask0.end());
            out.mparent.push_back(o.mpar);
            out.mmask.push_back(o.mmask1);
            out.mmask0.push_back(o.mmask0);
            out.mregex.push_back(o.mreg);
            if (out.eparent.size() > (1u << 26)) {
                return false;
            }
        }
    }
    out.eofs.push_back((uint32_t) out.eparent.size());
    out.nstates = (uint32_t) states.size();

    out.any_idx.assign(out.nstates, 0xff);
    out.eof_idx.assign(out.nstates, 0xff);
    out.eof_regex.assign(out.nstates, 0);
    out.eof_mask0.assign(out.nstates, 0);
    out.list_ofs.assign(out.nstates + 1, 0);
    for (uint32_t s = 0; s < out.nstates; s++) {
        const list_t &l = states[s].second;
        out.list_ofs[s] = (uint32_t) out.list_park.size();
        for (size_t j = 0; j < l.size(); j++) {
            const uint32_t P = l[j] & 0x7fff;
            out.list_park.push_back((uint16_t) P);
            if ((int32_t) P == T.p_any) {
                out.any_idx[s] = (uint8_t) j;
            }
        }
        /* the step at the end of the input: the first thread that reaches MATCH there */
        if (s != 0) {
            step_out_t o;
            step(states[s], true, 0, o);
            if (o.mev) {
                out.eof_idx[s] = o.mpar;
                out.eof_regex[s] = o.mreg;
                out.eof_mask0[s] = o.mmask0;
            }
        }
    }
    out.list_ofs[out.nstates] = (uint32_t) out.list_park.size();
    (void) prog;
    return true;
}
