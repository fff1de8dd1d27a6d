/*
 * sre_quirk.h -- on which inputs the reference's first-byte prefilter can misfire.
 *
 * sre_vm_pike_exec recognises "the thread list is the initial one" by the thread
 * COUNT and every pc but the LAST (sre_vm_pike.c:262-274).  After a match has cut
 * the ".*?" thread, the survivors of the step can pass that test; if the step was
 * the one a prefilter jump landed on (seen_start_state still set, :276-305), the
 * reference drops them, starts over at the next leading byte, and a later match
 * overwrites the leftmost one.  That takes: a landing byte on
 * which the fresh initial list reports a match at once, with survivors of the
 * initial list's length and leading pcs.  This analysis finds those bytes by
 * replaying add_thread (:756-942) on the bytecode; the batch Pike tiers re-run
 * the lines on which one of them starts the match in the faithful kernel
 * (kernels/sre_pike.cu: k_pike_quirk_mark, pike_exec(faithful)).
 */
#ifndef SRE_QUIRK_H
#define SRE_QUIRK_H

#include <stdint.h>
#include "../host/sre_internal.h"

/* -> true when the misfire is possible at all; single[b >> 5] bit b & 31: a match that starts at
 * byte value b right after a prefilter jump may trigger it.  Programs with assertions are
 * replayed in each look-behind context a landing offset can have (the byte in front a newline, a
 * word byte, neither), with the look-ahead threads parked and resolved by the landing byte as
 * the reference does (:452-509, :842-887): `\bGET\b`, `^\d+`, `\w+$` have an empty set. */
bool sre_quirk_bytes(const sre_program_t *prog, uint32_t single[8]);

#endif
