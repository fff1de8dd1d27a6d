/*
 * sre_pdfa.h -- the Pike VM determinised: a DFA over ORDERED thread lists, with
 * the provenance of every thread, for the capture kernel sre_pike_lineage.cu.
 *
 * New in this build.  What it replaces: the per-byte work of sre_vm_pike_exec
 * (sre_vm_pike.c:148-689) and sre_vm_pike_add_thread (:756-942).  The Pike VM's
 * thread list after a step is an ordered list of parked instructions (one
 * thread per instruction: the dedup of :770-792), and the next list is a
 * function of that list and the byte alone; captures only ride along.  So the
 * lists are the states of a DFA ("P-DFA"), built here with exactly the step
 * rule of the closure tables (sre_closure.h), and per transition the tables
 * record where every thread of the next list came from:
 *
 *   trans[s][c]      next state; bit 15: the step reports a match (a closure
 *                    reached MATCH, :889-899, or a parked MATCH thread was met)
 *                    and cuts the threads of lower priority (:535-553)
 *   eofs/eparent/emask   for thread j of the next list: the index of its
 *                    parent in list s and the capture slots SAVEd on the way
 *                    (they take the position after the consumed byte)
 *   mparent/mmask/mregex  the same for the thread that matched
 *
 * A line is then matched in two cheap passes: forward, one table look-up per
 * byte, remembering the state at each position and the last match event; then
 * backward along the lineage of the ONE thread that won, filling each capture
 * slot from the last step that SAVEd it.  Leftmost-first priority is the list
 * order, so the result is the Pike VM's, bit for bit.
 *
 * Assertions are part of it.  Look-behind ones (`\A`, `^`): the closure a step
 * appends depends on the consumed byte only through "was it a newline" (the
 * look-behind context of sre_closure.h), so the state remembers the kind of the
 * byte in front of it (nothing / newline / word byte / other, as far as the
 * program can tell them apart) and the search starts from one of four start
 * lists.  Look-ahead ones (`$ \z \b \B`): a thread parked on the assertion is
 * part of the list (with the seen_word flag of sre_vm_pike.c:868-887 for `\b
 * \B`); the step on the next byte -- or the step at the end of the input --
 * resolves it and splices its closure in at the SAME position, so a provenance
 * record carries two slot sets: SAVEd at the position of the step (emask0) and
 * after the consumed byte (emask).  sre_build_pdfa returns false when the
 * automaton exceeds max_states / lists of 255 threads: those programs stay on
 * k_pike_table.
 */
#ifndef SRE_PDFA_H
#define SRE_PDFA_H

#include <stdint.h>
#include <vector>
#include "sre_closure.h"

struct sre_pdfa_t {
    uint32_t                nstates = 0;        /* state 0 = the empty list           */
    uint32_t                nclasses = 0;
    uint8_t                 clsmap[256];
    /* the start closure by what lies in front of the first byte: nothing (offset 0), a newline, a
     * word byte, anything else */
    uint32_t                init[4] = { 0, 0, 0, 0 };
    uint32_t                init_mask_ofs[4] = { 0, 0, 0, 0 };  /* ... where its init_mask begins */
    bool                    ctx_dep = false;    /* program has \A or ^ */
    bool                    lookahead = false;  /* program has $ \z \b \B */
    uint32_t                max_slots = 0;      /* slots of the largest regex (<= 32) */
    std::vector<uint16_t>   trans;              /* [nstates][nclasses]                */
    std::vector<uint32_t>   eofs;               /* [nstates * nclasses + 1]           */
    std::vector<uint8_t>    eparent;
    std::vector<uint32_t>   emask, emask0;      /* slots SAVEd after the byte / at the step's position */
    std::vector<uint8_t>    mparent;            /* [nstates * nclasses]               */
    std::vector<uint32_t>   mmask, mmask0;
    std::vector<uint16_t>   mregex;
    std::vector<uint8_t>    any_idx;            /* [nstates] index of the ".*?" thread, 0xff: none */
    std::vector<uint8_t>    eof_idx;            /* [nstates] the thread that reaches MATCH in the step at
                                                   the end of the input (its index in the list), 0xff */
    std::vector<uint16_t>   eof_regex;          /* [nstates]                          */
    std::vector<uint32_t>   eof_mask0;          /* [nstates] slots SAVEd on the way (at the end position) */
    std::vector<uint32_t>   init_mask;          /* slots SAVEd by a start closure, per thread (init_mask_ofs) */
    std::vector<uint32_t>   list_ofs;           /* [nstates + 1] the lists themselves */
    std::vector<uint16_t>   list_park;
};

bool sre_build_pdfa(const sre_program_t *prog, const sre_closure_table_t &T, uint32_t max_states, sre_pdfa_t &out);

#endif
