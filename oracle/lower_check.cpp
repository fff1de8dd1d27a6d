/*
 * lower_check.cpp -- TEST INFRASTRUCTURE ONLY: a host executor over the
 * product's lowered tables (sregex_b200/csrc/lower), so that the lowering pass
 * can be validated on the whole corpus in the CPU-only test tier, before and
 * independently of any kernel.  It is never linked into libsregex_cuda and is
 * not a fallback: the product's exec entry points only launch CUDA kernels.
 *
 * It applies literally the step rule documented in sre_lower.h.
 */
#include <cstring>

#include "../sregex_b200/csrc/lower/sre_lower.h"

struct lc_t {
    sre_lowered_t           low;
    std::vector<uint32_t>   S;          /* streaming NFA state */
    bool                    started = false;
    uint32_t                dstate = 0; /* streaming DFA state */
};

extern "C" {

lc_t *lc_create(sre_program_t *prog, unsigned max_dfa_states)
{
    lc_t *lc = new lc_t();
    if (sre_lower_program(prog, max_dfa_states, &lc->low) != SRE_OK) {
        delete lc;
        return NULL;
    }
    return lc;
}

void lc_destroy(lc_t *lc) { delete lc; }

void lc_info(lc_t *lc, unsigned *out)
{
    out[0] = lc->low.nfa.nstates;
    out[1] = lc->low.nfa.nclasses;
    out[2] = lc->low.nfa.nkinds;
    out[3] = lc->low.has_dfa ? lc->low.dfa.nstates : 0;
    out[4] = lc->low.has_dfa ? lc->low.dfa.nclasses : 0;
    unsigned shift = 0;
    for (uint32_t s = 0; s < lc->low.nfa.nstates; s++) {
        shift += (lc->low.nfa.shift_mask[s >> 5] >> (s & 31)) & 1;
    }
    out[5] = shift;
}

void lc_reset(lc_t *lc)
{
    lc->started = false;
    lc->dstate = lc->low.dfa.start;
}

/* one sre_vm_thompson_exec()-shaped call over the lowered NFA */
long lc_nfa_exec(lc_t *lc, const uint8_t *buf, size_t len, unsigned eof)
{
    const sre_nfa_t &n = lc->low.nfa;
    const uint32_t W = n.nwords;
    if (!lc->started) {
        lc->started = true;
        lc->S = n.init;
    }
    std::vector<uint32_t> next(W);
    for (size_t i = 0; i < len; i++) {
        uint32_t c = n.clsmap[buf[i]];
        const uint32_t *mv = &n.mv[(size_t) c * W], *mt = &n.mt[(size_t) c * W];
        bool any = false;
        for (uint32_t w = 0; w < W; w++) {
            if (lc->S[w] & mt[w]) return SRE_OK;
            any |= lc->S[w] != 0;
        }
        if (!any) break;
        std::fill(next.begin(), next.end(), 0u);
        for (uint32_t w = 0; w < W; w++) {
            uint32_t m = lc->S[w] & mv[w];
            while (m) {
                uint32_t s = w * 32 + __builtin_ctz(m);
                m &= m - 1;
                const uint32_t *row = n.follow_row(n.cls_kind[c], s);
                for (uint32_t x = 0; x < W; x++) next[x] |= row[x];
            }
        }
        lc->S.swap(next);
    }
    if (eof) {
        for (uint32_t w = 0; w < W; w++) {
            if (lc->S[w] & n.mt_eof[w]) return SRE_OK;
        }
        return SRE_DECLINED;
    }
    return SRE_AGAIN;
}

/* same, over the DFA (byte-class table, and the 256-wide table if present) */
long lc_dfa_exec(lc_t *lc, const uint8_t *buf, size_t len, unsigned eof, int use_t256)
{
    const sre_dfa_t &d = lc->low.dfa;
    if (!lc->low.has_dfa) return SRE_ERROR;
    if (!lc->started) {
        lc->started = true;
        lc->dstate = d.start;
    }
    uint32_t s = lc->dstate;
    bool t256 = use_t256 && !d.t256.empty();
    for (size_t i = 0; i < len && s != d.acc; i++) {
        s = t256 ? d.t256[(size_t) s * 256 + buf[i]]
                 : d.trans[(size_t) s * d.nclasses + d.clsmap[buf[i]]];
    }
    lc->dstate = s;
    if (s == d.acc) return SRE_OK;
    if (eof) return d.fin[s] ? SRE_OK : SRE_DECLINED;
    return SRE_AGAIN;
}

}

/* transfer function of a buffer over the DFA: out[d] = state reached from d
 * (test stand-in for the GPU stream reduce in the world_size-2 gloo tests) */
extern "C" int lc_dfa_fn(lc_t *lc, const uint8_t *buf, size_t len, uint8_t *out)
{
    const sre_dfa_t &d = lc->low.dfa;
    if (!lc->low.has_dfa || d.nstates > 256) return SRE_ERROR;
    for (uint32_t e = 0; e < d.nstates; e++) {
        uint32_t s = e;
        for (size_t i = 0; i < len; i++) {
            s = d.trans[(size_t) s * d.nclasses + d.clsmap[buf[i]]];
        }
        out[e] = (uint8_t) s;
    }
    return SRE_OK;
}

extern "C" int lc_dfa_fin(lc_t *lc, unsigned state)
{
    return lc->low.has_dfa && state < lc->low.dfa.nstates ? lc->low.dfa.fin[state] : 0;
}

/* Pike start hint over the restart table (see sre_lower.h h256); -1 if the
 * program has no such table */
extern "C" long lc_hint(lc_t *lc, const uint8_t *buf, size_t len)
{
    const sre_dfa_t &d = lc->low.dfa;
    if (!lc->low.has_dfa || d.h256.empty()) return -1;
    uint32_t s = 0;
    long p0 = 0;
    for (size_t i = 0; i < len; i++) {
        s = d.h256[(size_t) s * 256 + buf[i]];
        if (s & 0x80) p0 = (long) i + 1;
    }
    return p0;
}

/* the same hint from the class-compressed restart table (any DFA size) */
extern "C" long lc_hint_cls(lc_t *lc, const uint8_t *buf, size_t len)
{
    const sre_dfa_t &d = lc->low.dfa;
    if (!lc->low.has_dfa || d.hcls.empty()) return -1;
    uint32_t s = d.start;
    long p0 = 0;
    for (size_t i = 0; i < len && s != d.acc; i++) {
        const uint16_t e = d.hcls[(size_t) s * d.hncls + d.hclsmap[buf[i]]];
        s = e & 0x7fff;
        if (e & 0x8000) p0 = (long) i + 1;
    }
    return p0;
}

/*
 * CPU model of k_pike_table (sregex_b200/csrc/kernels/sre_pike_table.cu): the
 * Pike VM driven by the closure tables of sre_closure.h, for one buffer with
 * eof = 1 on a fresh context, started at offset `start`.  Thread lists are
 * unbounded here (the kernel reports lines beyond its capacities as RETRY).
 * Returns the rc of sre_vm_pike_exec, or -1000 when the program has no table;
 * ovec receives 2 * (ncaps + 1) values.
 */
#include "../sregex_b200/csrc/lower/sre_closure.h"

namespace {

struct tp_thread_t {
    uint32_t             park;
    bool                 seen_word;
    std::vector<int64_t> cap;
};

inline bool tp_isword(uint32_t c)
{
    return (c - '0' < 10u) || ((c | 0x20) - 'a' < 26u) || c == '_';
}

}  // namespace

extern "C" long lc_table_pike(sre_program_t *prog, const uint8_t *input, long size, long start, int64_t *ovec)
{
    sre_closure_table_t T;
    if (!sre_build_closure_table(prog, 4096, T)) {
        return -1000;
    }
    const uint32_t np = T.npark;
    const size_t nslots = T.max_slots;          /* a thread carries the slots of its own regex */
    /* dedup marks: "marked in this step" / "marked in the previous step" */
    std::vector<uint8_t> m_cur(np, 0), m_prev(np, 0);
    std::vector<tp_thread_t> clist, nlist, hold;     /* hold: LIFO, back = top */
    std::vector<int64_t> matched_cap;
    long matched_id = 0;
    bool matched = false;

    /* 0 ok, 1 MATCH reached (want_done) */
    auto append_closure = [&](uint32_t P, long pos, const std::vector<int64_t> &parent,
                              std::vector<tp_thread_t> &out, bool use_prev, bool want_done) -> int {
        uint32_t v = 0, prev = 0;
        if (pos > 0) {
            prev = input[pos - 1];
            v = T.ctx_dep ? (prev == '\n' ? 1u : 2u) : 0u;
        }
        const int nb = pos < size ? (int) input[pos] : -2;
        const uint32_t *list = T.ent.data(), *lmask = T.emask.data();
        uint32_t e0 = T.ofs[v * (np + 2) + P], e1 = T.ofs[v * (np + 2) + P + 1];
        bool filtered = false;
        if ((P == np || (int32_t) P == T.p_any) && !T.bent.empty() && nb >= 0) {
            list = T.bent.data();           /* the start closure by next byte */
            lmask = T.bmask.data();
            e0 = T.bofs[v * 257 + nb];
            e1 = T.bofs[v * 257 + nb + 1];
            filtered = true;
        }
        for (uint32_t e = e0; e < e1; e++) {
            const uint32_t fp = list[e], mask = lmask[e];
            const uint32_t kind = T.kind[fp];
            if (!filtered && kind == 0
                && (nb == -2 || !((T.accept[T.acc_idx[fp] * 8 + ((uint32_t) nb >> 5)] >> (nb & 31)) & 1)))
            {
                continue;
            }
            std::vector<uint8_t> &marks = use_prev ? m_prev : m_cur, &other = use_prev ? m_cur : m_prev;
            if (marks[fp]) {
                continue;
            }
            marks[fp] = 1;
            other[fp] = 0;
            std::vector<int64_t> cap = parent;
            for (size_t s = 0; s < nslots; s++) {
                if ((mask >> s) & 1) {
                    cap[s] = pos;
                }
            }
            if (kind == 1 && want_done) {
                matched_cap = cap;
                matched_id = T.regex[fp];
                return 1;
            }
            tp_thread_t t;
            t.park = fp;
            t.seen_word = kind >= 4 && pos > 0 && tp_isword(prev);
            t.cap = cap;
            out.push_back(t);
        }
        return 0;
    };

    const std::vector<int64_t> none(nslots, -1);
    long sp = start;
    append_closure(np, sp, none, clist, false, false);
    for (; sp <= size; sp++) {
        if (clist.empty()) {
            break;
        }
        m_prev = m_cur;
        std::fill(m_cur.begin(), m_cur.end(), 0);
        const bool at_end = (sp == size);
        const uint32_t byte = at_end ? 0 : input[sp];
        const bool cur_word = !at_end && tp_isword(byte);
        size_t i = 0;
        hold.clear();
        for (;;) {
            tp_thread_t t;
            if (!hold.empty()) {
                t = hold.back();
                hold.pop_back();
            } else if (i < clist.size()) {
                t = clist[i++];
            } else {
                break;
            }
            const uint32_t kind = T.kind[t.park];
            bool got_match = false;
            if (kind >= 2) {
                bool holds;
                switch (kind) {
                case 2:  holds = at_end; break;
                case 3:  holds = at_end || byte == '\n'; break;
                case 4:  holds = (t.seen_word == cur_word); break;
                default: holds = (t.seen_word != cur_word); break;
                }
                if (holds) {
                    std::vector<tp_thread_t> add;
                    append_closure(t.park, sp, t.cap, add, true, false);
                    for (size_t k = add.size(); k-- > 0;) {     /* prepended: first one on top */
                        hold.push_back(add[k]);
                    }
                }
            } else if (kind == 1) {
                matched_cap = t.cap;
                matched_id = T.regex[t.park];
                got_match = true;
            } else if (!at_end && ((T.accept[T.acc_idx[t.park] * 8 + (byte >> 5)] >> (byte & 31)) & 1)) {
                got_match = append_closure(t.park, sp + 1, t.cap, nlist, false, true) == 1;
            }
            if (got_match) {
                matched = true;
                break;
            }
        }
        clist.swap(nlist);
        nlist.clear();
        if (at_end) {
            break;
        }
    }
    if (!matched) {
        for (size_t s = 0; s < nslots; s++) {
            ovec[s] = -1;
        }
        return SRE_DECLINED;
    }
    /* prepare_matched_captures :945-989: the matched regex's slots, then -1 */
    const size_t cnt = 2 * (size_t) (prog->multi_ncaps[matched_id] + 1);
    for (size_t s = 0; s < nslots; s++) {
        ovec[s] = s < cnt ? matched_cap[s] : -1;
    }
    return matched_id;
}

/*
 * Host model of the candidate-based stream scan of kernels/sre_stream.cu (see
 * sre_image.h): per piece the image automaton walks the `window` bytes before
 * the piece and yields the candidate entry states; the piece is run from each;
 * the pieces are then chained by looking the carried state up among the
 * candidates.  A piece whose window stays wide is "unresolved" and run from the
 * carried state itself.  Checks the claim the kernels rest on -- the carried
 * state is always among the candidates -- on the CPU:
 * returns 0, or -2 when a carried state was not among a piece's candidates.
 * out[0] = exit state, out[1] = offset of the step entering ACC (-1: none),
 * out[2] = unresolved pieces, out[3] = largest candidate count seen.
 */
#include "../sregex_b200/csrc/lower/sre_image.h"

extern "C" int lc_stream_model(lc_t *lc, const uint8_t *buf, size_t len, size_t piece, size_t window,
    unsigned K, unsigned entry, long long *out)
{
    const sre_dfa_t &d = lc->low.dfa;
    if (!lc->low.has_dfa) return SRE_ERROR;
    sre_image_t U;
    if (!sre_build_image_automaton(d, K, 60000, 512, U)) return SRE_ERROR;
    const uint32_t C = d.nclasses;
    auto step = [&](uint32_t s, uint8_t b) -> uint32_t {
        return s == d.acc ? s : d.trans[(size_t) s * C + d.clsmap[b]];
    };
    uint32_t s = entry;
    long long first = -1, unresolved = 0, maxc = 0;
    for (size_t p = 0; p < len || p == 0; p += piece) {
        const size_t e = p + piece < len ? p + piece : len;
        std::vector<uint32_t> cand;
        bool resolved = true;
        if (p == 0) {
            cand.push_back(entry);
        } else {
            resolved = false;
            for (size_t w = window; !resolved; w *= 16) {
                const size_t from = w < p ? p - w : 0;
                uint32_t u = 0;
                for (size_t i = from; i < p; i++) {
                    u = U.trans[(size_t) u * C + d.clsmap[buf[i]]];
                }
                if (U.ncand[u] != 0xff) {
                    resolved = true;
                    for (uint32_t j = 0; j < U.ncand[u]; j++) cand.push_back(U.cand[(size_t) u * K + j]);
                }
                if (from == 0 || w >= piece) break;
            }
        }
        if ((long long) cand.size() > maxc) maxc = (long long) cand.size();
        uint32_t t = s;
        if (s == d.acc) {
            /* absorbed earlier */
        } else if (!resolved) {
            unresolved++;
            for (size_t i = p; i < e; i++) t = step(t, buf[i]);
        } else {
            bool found = false;
            for (uint32_t c0 : cand) {
                if (c0 == s) found = true;
            }
            if (!found) return -2;
            for (size_t i = p; i < e; i++) t = step(t, buf[i]);
        }
        if (t == d.acc && s != d.acc && first < 0) {
            uint32_t x = s;
            for (size_t i = p; i < e; i++) {
                x = step(x, buf[i]);
                if (x == d.acc) { first = (long long) i; break; }
            }
        }
        s = t;
        if (len == 0) break;
    }
    out[0] = s;
    out[1] = first;
    out[2] = unresolved;
    out[3] = maxc;
    return 0;
}

/* image automaton statistics: out[0] = U-states, out[1] = narrow ones */
extern "C" int lc_image_info(lc_t *lc, unsigned K, unsigned *out)
{
    sre_image_t U;
    if (!lc->low.has_dfa || !sre_build_image_automaton(lc->low.dfa, K, 60000, 512, U)) return SRE_ERROR;
    out[0] = U.nstates;
    out[1] = 0;
    for (uint32_t u = 0; u < U.nstates; u++) out[1] += U.ncand[u] != 0xff;
    return SRE_OK;
}

/*
 * CPU stand-in for sre_cuda_thompson_stream_reduce in the world_size-2 gloo
 * tests: the record of a whole stream part in the format of kernels/sre_stream.cu
 * (8 candidate entry states as u16, then the 8 states they lead to; 0xffff =
 * empty slot; first candidate 0xfffe = unresolved).  Candidates: {entry} when
 * known (entry != 0xfffffffe), else the image of all states under the halo.
 */
extern "C" int lc_stream_part_record(lc_t *lc, const uint8_t *buf, size_t len, const uint8_t *halo, size_t halo_len,
    unsigned entry, uint8_t *out32)
{
    const sre_dfa_t &d = lc->low.dfa;
    const unsigned K = 8;
    if (!lc->low.has_dfa) return SRE_ERROR;
    uint16_t rec[16];
    for (int i = 0; i < 16; i++) rec[i] = 0xffff;
    std::vector<uint32_t> cand;
    if (entry != 0xfffffffeu) {
        if (entry != d.acc) cand.push_back(entry);
    } else if (halo == NULL) {
        rec[0] = 0xfffe;
        memcpy(out32, rec, 32);
        return SRE_OK;
    } else {
        sre_image_t U;
        if (!sre_build_image_automaton(d, K, 60000, 512, U)) return SRE_ERROR;
        uint32_t u = 0;
        for (size_t i = 0; i < halo_len; i++) u = U.trans[(size_t) u * d.nclasses + d.clsmap[halo[i]]];
        if (U.ncand[u] == 0xff) {
            rec[0] = 0xfffe;
            memcpy(out32, rec, 32);
            return SRE_OK;
        }
        for (uint32_t j = 0; j < U.ncand[u]; j++) cand.push_back(U.cand[(size_t) u * K + j]);
    }
    for (size_t j = 0; j < cand.size(); j++) {
        uint32_t s = cand[j];
        for (size_t i = 0; i < len && s != d.acc; i++) s = d.trans[(size_t) s * d.nclasses + d.clsmap[buf[i]]];
        rec[j] = (uint16_t) cand[j];
        rec[8 + j] = (uint16_t) s;
    }
    memcpy(out32, rec, 32);
    return SRE_OK;
}

/* exit state and offset of the first step entering ACC (-1) of a part run from `entry` */
extern "C" long long lc_stream_part_run(lc_t *lc, const uint8_t *buf, size_t len, unsigned entry, unsigned *exit_state)
{
    const sre_dfa_t &d = lc->low.dfa;
    uint32_t s = entry;
    long long first = -1;
    for (size_t i = 0; i < len && s != d.acc; i++) {
        s = d.trans[(size_t) s * d.nclasses + d.clsmap[buf[i]]];
        if (s == d.acc) first = (long long) i;
    }
    *exit_state = s;
    return first;
}

/*
 * Host model of kernels/sre_pike_lineage.cu over the P-DFA (sre_pdfa.h): the
 * forward pass (one table look-up per byte, the state of the last `ring`
 * positions remembered, the last match event noted), then the backward walk
 * along the winning thread's lineage.  Returns the matched regex id /
 * SRE_DECLINED like lc_table_pike, -1000 when the program has no P-DFA, -1001
 * when the lineage is older than the ring (the kernel hands such a line to the
 * next tier).  out_info (may be NULL): [0] = P-DFA states, [1] = classes.
 */
#include "../sregex_b200/csrc/lower/sre_pdfa.h"

extern "C" long lc_pdfa_pike(sre_program_t *prog, const uint8_t *input, long size, long start, long ring,
    int64_t *ovec, unsigned *out_info)
{
    sre_closure_table_t T;
    sre_pdfa_t D;
    if (!sre_build_closure_table(prog, 4096, T) || !sre_build_pdfa(prog, T, 4096, D)) {
        return -1000;
    }
    if (out_info) {
        out_info[0] = D.nstates;
        out_info[1] = D.nclasses;
    }
    const uint32_t C = D.nclasses;
    const size_t nslots = T.max_slots;
    std::vector<uint16_t> hist((size_t) ring, 0);
    /* the start list by what lies in front of the first byte: nothing / '\n' / a word byte / other */
    const uint32_t v0 = start <= 0 ? 0u : input[start - 1] == '\n' ? 1u : tp_isword(input[start - 1]) ? 2u : 3u;
    uint32_t s = D.init[v0];
    long pos = start, mpos = -1;
    bool have = false, at_eof = false;
    for (; pos < size && s != 0; pos++) {
        const uint32_t c = D.clsmap[input[pos]];
        hist[(size_t) (pos % ring)] = (uint16_t) s;
        const uint16_t e = D.trans[(size_t) s * C + c];
        if (e & 0x8000) {
            have = true;
            mpos = pos;
        }
        s = e & 0x7fff;
    }
    /* the step at the end of the input (only when the input was walked to its end) */
    if (s != 0 && pos == size && D.eof_idx[s] != 0xff) {
        have = true;
        at_eof = true;
    }
    for (size_t k = 0; k < nslots; k++) {
        ovec[k] = -1;
    }
    if (!have) {
        return SRE_DECLINED;
    }
    long oldest = pos - ring;               /* positions > oldest are still in the ring */
    /* positions the ring has lost: forward again from the start, up to and including `upto`
     * (k_pike_lineage's refill; at most 64 times per input, then the next tier) */
    int refills = 0;
    std::vector<uint16_t> ckpt;
    auto refill = [&](long upto) -> bool {
        if (refills == 64) {
            return false;
        }
        uint32_t r = D.init[v0];
        long p = start;
        if (refills > 0) {
            /* from the last noted state that still covers the `ring` positions up to `upto` */
            long i = (upto - (ring - 1) - start) / ring;
            i = upto - (ring - 1) < start ? 0 : i >= (long) ckpt.size() ? (long) ckpt.size() - 1 : i;
            r = ckpt[(size_t) i];
            p = start + i * ring;
        }
        const bool note = refills == 0;
        refills++;
        for (; p <= upto; p++) {
            if (note && (p - start) % ring == 0 && ckpt.size() < 32) {
                ckpt.push_back((uint16_t) r);       /* the state in front of every ring-th position */
            }
            hist[(size_t) (p % ring)] = (uint16_t) r;
            r = D.trans[(size_t) r * C + D.clsmap[input[p]]] & 0x7fff;
        }
        oldest = upto - ring;
        return true;
    };
    uint32_t cur, j, rid;
    long u;
    uint32_t unset = nslots >= 32 ? 0xffffffffu : ((1u << nslots) - 1);
    auto assign = [&](uint32_t mask, long where) {
        uint32_t m = mask & unset;
        unset &= ~mask;
        for (size_t k = 0; k < nslots; k++) {
            if ((m >> k) & 1) ovec[k] = where;
        }
    };
    if (at_eof) {
        cur = s;
        j = D.eof_idx[s];
        rid = D.eof_regex[s];
        assign(D.eof_mask0[s], size);
        u = size - 1;
    } else {
        if (mpos <= oldest && !refill(mpos)) return -1001;
        cur = hist[(size_t) (mpos % ring)];
        const size_t t = (size_t) cur * C + D.clsmap[input[mpos]];
        j = D.mparent[t];
        rid = D.mregex[t];
        assign(D.mmask[t], mpos + 1);
        assign(D.mmask0[t], mpos);
        u = mpos - 1;
    }
    /* j indexes the list of state `cur`, the state before step u + 1 */
    for (;;) {
        if (j == D.any_idx[cur]) {
            break;                          /* the ".*?" thread carries no captures */
        }
        if (u < start) {
            assign(D.init_mask[D.init_mask_ofs[v0] + j], start);    /* a thread of the start closure */
            break;
        }
        if (u <= oldest && !refill(u)) return -1001;
        const uint32_t before = hist[(size_t) (u % ring)];
        const size_t idx = D.eofs[(size_t) before * C + D.clsmap[input[u]]] + j;
        assign(D.emask[idx], u + 1);
        assign(D.emask0[idx], u);
        j = D.eparent[idx];
        cur = before;
        u--;
    }
    const size_t cnt = 2 * (size_t) (prog->multi_ncaps[rid] + 1);
    for (size_t k = cnt; k < nslots; k++) {
        ovec[k] = -1;
    }
    return (long) rid;
}

/*
 * The bytes on which the reference's first-byte prefilter can misfire
 * (lower/sre_quirk.h) and the leading-byte set, for the CPU test of the marking
 * rule of k_pike_quirk_mark: single / lead = 256-bit sets; -> 1 when the misfire
 * is possible at all for this program.
 */
#include "../sregex_b200/csrc/lower/sre_quirk.h"

extern "C" int lc_quirk_bytes(sre_program_t *prog, uint32_t *single, uint32_t *lead)
{
    memset(lead, 0, 32);
    for (uint32_t b = 0; b < 256; b++) {
        for (uint32_t i = 0; i < prog->nleading; i++) {
            const sre_instruction_t &in = prog->insts[prog->leading[i]];
            bool t = false;
            if (in.opcode == SRE_OPCODE_CHAR) {
                t = b == in.ch;
            } else if (in.opcode == SRE_OPCODE_ANY) {
                t = true;
            } else if (in.opcode == SRE_OPCODE_IN || in.opcode == SRE_OPCODE_NOTIN) {
                bool inr = false;
                for (uint32_t j = 0; j < in.nranges; j++) {
                    inr |= b >= prog->ranges[in.v + j].from && b <= prog->ranges[in.v + j].to;
                }
                t = inr == (in.opcode == SRE_OPCODE_IN);
            }
            if (t) {
                lead[b >> 5] |= 1u << (b & 31);
            }
        }
    }
    return sre_quirk_bytes(prog, single) ? 1 : 0;
}

/*
 * The fast Pike tiers (closure tables, P-DFA) deduplicate closures by parked instruction; the
 * reference does it by one tag word per instruction (:770-792), and a look-ahead thread that
 * holds appends its closure under the tag of the PREVIOUS step (:484-509).  The two agree unless
 * an instruction can be visited both by such a held closure and by an ordinary one.
 * -> 0: no look-ahead assertion, or nothing behind one is visited by an ordinary closure;
 *    1: some instruction is visited both ways, but never by one closure that also parks the
 *       assertion (`foo\b|bar`: the tail after the alternation);
 *    2: one closure parks a look-ahead assertion and visits what lies behind it (an assertion
 *       that can be skipped: `(\B)?x`, `(?:\b|x)+`): the one program shape on which the
 *       batch Pike tiers are known to differ from the reference.
 */

/*
 * reach(root): the instructions add_thread visits from `root` without consuming (JMP / SPLIT /
 * SAVE / look-behind assertions passed through; consuming instructions, MATCH and look-ahead
 * assertions visited and not passed).
 */
namespace {

void reach(const sre_program_t *prog, int32_t pc, std::vector<uint8_t> &seen)
{
    if (pc < 0 || (uint32_t) pc >= prog->len || seen[pc]) {
        return;
    }
    seen[pc] = 1;
    const sre_instruction_t &in = prog->insts[pc];
    switch (in.opcode) {
    case SRE_OPCODE_JMP:
        reach(prog, in.x, seen);
        break;
    case SRE_OPCODE_SPLIT:
        reach(prog, in.x, seen);
        reach(prog, in.y, seen);
        break;
    case SRE_OPCODE_SAVE:
        reach(prog, pc + 1, seen);
        break;
    case SRE_OPCODE_ASSERT:
        if (in.v == SRE_REGEX_ASSERT_BIG_A || in.v == SRE_REGEX_ASSERT_CARET) {
            reach(prog, pc + 1, seen);
        }
        break;
    default:
        break;
    }
}

inline bool is_lookahead(const sre_instruction_t &in)
{
    return in.opcode == SRE_OPCODE_ASSERT && in.v != SRE_REGEX_ASSERT_BIG_A && in.v != SRE_REGEX_ASSERT_CARET;
}

inline bool consumes(const sre_instruction_t &in)
{
    return in.opcode == SRE_OPCODE_CHAR || in.opcode == SRE_OPCODE_ANY || in.opcode == SRE_OPCODE_IN
           || in.opcode == SRE_OPCODE_NOTIN;
}

}  // namespace

static int sre_lookahead_overlap(const sre_program_t *prog)
{
    const uint32_t n = prog->len;
    /* what the closures behind the look-ahead assertions visit (held closures, transitively) */
    std::vector<std::vector<uint8_t> > held(n);
    bool any = false;
    for (uint32_t a = 0; a < n; a++) {
        if (is_lookahead(prog->insts[a])) {
            held[a].assign(n, 0);
            reach(prog, (int32_t) a + 1, held[a]);
            any = true;
        }
    }
    if (!any) {
        return 0;
    }
    /* a held closure can park another look-ahead assertion, which is resolved in the same step */
    for (bool grown = true; grown;) {
        grown = false;
        for (uint32_t a = 0; a < n; a++) {
            if (held[a].empty()) {
                continue;
            }
            for (uint32_t b = 0; b < n; b++) {
                if (b != a && held[a][b] && !held[b].empty()) {
                    for (uint32_t k = 0; k < n; k++) {
                        if (held[b][k] && !held[a][k]) {
                            held[a][k] = 1;
                            grown = true;
                        }
                    }
                }
            }
        }
    }
    int level = 0;
    std::vector<uint8_t> all_normal(n, 0), seen;
    for (int32_t root = -1; root < (int32_t) n; root++) {
        if (root >= 0 && !consumes(prog->insts[root])) {
            continue;
        }
        /* the closure a step appends: of the start (root -1) or behind a consuming instruction */
        seen.assign(n, 0);
        reach(prog, root + 1, seen);
        for (uint32_t k = 0; k < n; k++) {
            all_normal[k] |= seen[k];
        }
        /* one closure that parks a look-ahead assertion AND visits what lies behind it */
        for (uint32_t a = 0; a < n; a++) {
            if (!held[a].empty() && seen[a]) {
                for (uint32_t k = 0; k < n; k++) {
                    if (held[a][k] && seen[k]) {
                        return 2;
                    }
                }
            }
        }
    }
    /* anything visited both by a held closure and by an ordinary one */
    for (uint32_t a = 0; a < n; a++) {
        for (uint32_t k = 0; k < n && !held[a].empty(); k++) {
            if (held[a][k] && all_normal[k]) {
                level = 1;
            }
        }
    }
    return level;
}

extern "C" int lc_lookahead_overlap(sre_program_t *prog)
{
    return sre_lookahead_overlap(prog);
}
