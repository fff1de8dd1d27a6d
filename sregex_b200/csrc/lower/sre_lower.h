/*
 * sre_lower.h -- lowering pass: sre_program_t -> GPU tables.
 *
 * New in this build (the reference has no counterpart; the closest thing is
 * the analysis its x86-64 JIT performs at compile time,
 * sre_vm_thompson_x64.dasc:323-394 build_paths and :623-726 get_next_states,
 * which pre-computes, per consuming instruction, the epsilon closure together
 * with the set of assertions crossed).  See DESIGN.md section 3.
 *
 * Lowered NFA
 *   state  = (pc, allow): pc is a consuming instruction (CHAR/ANY/IN/NOTIN) or
 *            a MATCH; `allow` is the set of *next-byte kinds* under which the
 *            thread survives its pending look-ahead assertions ($ \z \b \B):
 *            bit 0 EOF, bit 1 '\n', bit 2 word byte, bit 3 other byte.
 *   follow[k][s] = bitset of states entered after state s consumed a byte of
 *            kind k (k: 0 '\n', 1 word, 2 other); look-behind assertions
 *            (^ \A and the "previous byte" half of \b \B) are resolved here,
 *            at closure time, when the consumed byte is known.
 *   per byte class c:  mv[c] = states that survive AND consume a byte of class
 *            c;  mt[c] = MATCH states that survive a byte of class c.
 *   One step on byte b (class c, kind k):
 *            if (S & mt[c]) -> match seen at this step (sre_vm_thompson.c:233)
 *            S' = OR over s in (S & mv[c]) of follow[k][s]
 *   EOF step (the reference's extra iteration at sp == last when eof,
 *            sre_vm_thompson.c:88): match iff (S & mt_eof).
 *
 * DFA (optional, when the subset construction stays small): state 0.. with an
 * absorbing ACC state entered at the step where the NFA sees a match, plus a
 * per-state `fin` flag = "the EOF step sees a match".
 */
#ifndef SRE_LOWER_H
#define SRE_LOWER_H

#include <stdint.h>
#include <vector>
#include "../host/sre_internal.h"

enum { SRE_KIND_EOF = 0, SRE_KIND_NL = 1, SRE_KIND_WORD = 2, SRE_KIND_OTHER = 3 };
/* "previous position" kinds used while computing closures */
enum { SRE_PREV_NL = 0, SRE_PREV_WORD = 1, SRE_PREV_OTHER = 2, SRE_PREV_START = 3 };

struct sre_nfa_t {
    uint32_t                nstates = 0;
    uint32_t                nwords = 0;         /* 32-bit words per bitset      */
    std::vector<int32_t>    state_pc;
    std::vector<uint8_t>    state_allow;
    uint32_t                nclasses = 0;
    uint8_t                 clsmap[256];
    std::vector<uint8_t>    cls_kind;           /* [nclasses] 0..2              */
    uint32_t                nkinds = 1;         /* 1: follow is kind-blind      */
    std::vector<uint32_t>   mv;                 /* [nclasses][nwords]           */
    std::vector<uint32_t>   mt;                 /* [nclasses][nwords]           */
    std::vector<uint32_t>   mt_eof;             /* [nwords]                     */
    std::vector<uint32_t>   follow;             /* [nkinds][nstates][nwords]    */
    std::vector<uint32_t>   init;               /* [nwords], S at offset 0      */
    /* states whose follow set is exactly {s+1} for every kind: handled by a
     * shift instead of a row OR in the bit-parallel kernel */
    std::vector<uint32_t>   shift_mask;         /* [nwords]                     */
    bool                    has_match_lookahead = false;

    const uint32_t *follow_row(uint32_t kind, uint32_t s) const {
        return &follow[((size_t) (nkinds == 1 ? 0 : kind) * nstates + s) * nwords];
    }
};

struct sre_dfa_t {
    uint32_t                nstates = 0;        /* incl. ACC                    */
    uint32_t                start = 0;
    uint32_t                acc = 0;
    uint32_t                nclasses = 0;
    uint8_t                 clsmap[256];
    std::vector<uint16_t>   trans;              /* [nstates][nclasses]          */
    std::vector<uint8_t>    fin;                /* [nstates]                    */
    /* byte-indexed u8 table for the fast kernel, only when nstates <= 256    */
    std::vector<uint8_t>    t256;               /* [nstates][256]               */
    /*
     * "restart" table for the Pike start hint, only when nstates <= 128:
     * [256 rows][256]; row r and row r+128 are identical; entry = next state |
     * 0x80 when the only thread that consumed the byte is the ".*?" `any`
     * thread (every partial match died on it), so that a leftmost-first search
     * may be restarted right after this byte (Pike hint).
     */
    std::vector<uint8_t>    h256;
    /* the same information for any DFA size: [nstates][hncls] u16 over the NFA
     * byte classes (hclsmap), entry = next state | 0x8000 restart flag         */
    uint32_t                hncls = 0;
    uint8_t                 hclsmap[256];
    std::vector<uint16_t>   hcls;
};

struct sre_lowered_t {
    sre_nfa_t   nfa;
    bool        has_dfa = false;
    sre_dfa_t   dfa;
};

/* max_dfa_states: subset construction gives up beyond this (0 = no DFA) */
int sre_lower_program(const sre_program_t *prog, uint32_t max_dfa_states,
    sre_lowered_t *out);

#endif
