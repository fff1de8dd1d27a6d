/*
 * lower_check.cpp -- TEST INFRASTRUCTURE ONLY: a host executor over the
 * product's lowered tables (sregex_b200/csrc/lower), so that the lowering pass
 * can be validated on the whole corpus in the CPU-only test tier, before and
 * independently of any kernel.  It is never linked into libsregex_cuda and is
 * not a fallback: the product's exec entry points only launch CUDA kernels.
 *
 * It applies literally the step rule documented in sre_lower.h.
 */
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
