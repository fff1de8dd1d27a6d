/*
 * sre_image.h -- "image automaton" of the lowered DFA, for the chunk-parallel
 * stream scan (kernels/sre_stream.cu).
 *
 * New in this build.  What it serves: a sequence of sre_vm_thompson_exec calls
 * over one stream carries the thread lists from chunk to chunk
 * (sre_vm_thompson.c:63-270, ctx sre_vm_thompson.h:30-41); after lowering that
 * carry is ONE DFA state.  To scan pieces of the stream in parallel, a piece
 * must be summarised without knowing its entry state.  An N-state DFA piece is a
 * function on N states -- too big for N in the thousands.  But whatever the
 * state was L bytes before the piece, after those L bytes it can only be in the
 * IMAGE of all states under that window, and for regex automata on text that
 * image is tiny (a byte that no partial match survives leaves {start}).
 *
 * The image automaton U tracks that image over a window, as a DFA of its own:
 *   U-state 0 (TOP)  = nothing known: every DFA state possible;
 *   U-state u        = a set of DFA states (ACC left out: it is absorbing);
 *   step(u, class c) = { T[s][c] : s in set(u) } \ {ACC}.
 * Sets of at most K states are "narrow": their members are the candidate entry
 * states handed to the piece kernel.  Larger sets are kept exactly while the
 * budget lasts (so that a window such as /.{5}/ needs narrows by itself) and
 * are replaced by TOP beyond it -- a superset, which keeps every result exact:
 * the piece kernel only needs a set that CONTAINS the true entry state.
 */
#ifndef SRE_IMAGE_H
#define SRE_IMAGE_H

#include <stdint.h>
#include <vector>
#include "sre_lower.h"

struct sre_image_t {
    uint32_t                nstates = 0;    /* U-states; 0 = TOP                       */
    uint32_t                nclasses = 0;   /* == dfa.nclasses (dfa.clsmap applies)    */
    uint32_t                K = 0;          /* candidates per narrow state             */
    std::vector<uint16_t>   trans;          /* [nstates][nclasses] next U-state        */
    std::vector<uint16_t>   cand;           /* [nstates][K] DFA states, 0xffff padded  */
    std::vector<uint8_t>    ncand;          /* [nstates] 0..K, or 0xff: wide           */
};

/* K <= 16.  max_narrow / max_wide: budgets of numbered sets (beyond: TOP). */
bool sre_build_image_automaton(const sre_dfa_t &dfa, uint32_t K, uint32_t max_narrow,
    uint32_t max_wide, sre_image_t &out);

#endif
