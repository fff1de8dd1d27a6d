/*
 * sre_closure.h -- precomputed add_thread closures for the Pike kernels.
 *
 * New in this build.  The reference walks add_thread (sre_vm_pike.c:756-942)
 * at run time for every thread and every byte; its Thompson JIT precomputes
 * the closure of every consuming instruction instead
 * (sre_vm_thompson_x64.dasc:323-394).  These tables do the same for the Pike
 * VM, captures included; sre_pike_table.cu and sre_pike.cu consume them, and
 * oracle/lower_check.cpp runs them on the CPU against the oracle (test tier).
 */
#ifndef SRE_CLOSURE_H
#define SRE_CLOSURE_H

#include <stdint.h>
#include <vector>
#include "../host/sre_internal.h"

/* one thread of the start closure (== sre_dev_start_t of the kernels) */
struct sre_start_ent_t {
    int32_t   pc;
    uint8_t   nsl;          /* slots set to the current position ...          */
    uint8_t   sl[7];        /* ... relative to the owning regex's first slot  */
};

/*
 * Start closure by next byte (any program whose start closure is context
 * free): ents[ofs[b] .. ofs[b+1]) = the threads add_thread(pc 0) parks on
 * consuming instructions that can take byte b, in priority order.
 * pc_regex[pc] = regex owning pc, slot_ofs[r] = first capture slot of regex r.
 */
bool sre_build_start_closure(const sre_program_t *prog, const std::vector<uint16_t> &pc_regex,
    const std::vector<uint32_t> &slot_ofs, std::vector<uint32_t> &ofs, std::vector<sre_start_ent_t> &ents);

/*
 * Closure tables of a program (one regex or a set).  The instructions a thread
 * can be parked on (consuming, look-ahead assertion, MATCH) are numbered
 * 0 .. npark-1 in pc order.  Closure (ctx, P) = ent[ofs[ctx * (npark + 2) + P]
 * .. ofs[ctx * (npark + 2) + P + 1]) = what add_thread(pc(P) + 1) appends when
 * run on its own (P == npark: add_thread(0)), with `\A` and `^` decided by the
 * look-behind context ctx: 0 = at offset 0, 1 = after a newline, 2 = elsewhere.
 * Entry i = parked number ent[i] + the slots SAVEd on the path as a bit set
 * emask[i], the slots counted from the first slot of the regex that owns the
 * parked instruction (a thread only ever carries the slots of its own regex;
 * at most 32 of them, i.e. 15 groups).
 *
 * The start closure (P == npark, and P == p_any, the ".*?" thread, whose
 * closure is the same list) is also given bucketed by the next byte when it is
 * long: bent[bofs[ctx * 257 + b] .. bofs[ctx * 257 + b + 1]) = its entries that
 * are not consuming instructions unable to take byte b.
 */
struct sre_closure_table_t {
    std::vector<uint32_t> ent, emask;
    std::vector<uint16_t> ofs;
    std::vector<uint32_t> accept;       /* [nsets][8]: distinct byte sets */
    std::vector<uint16_t> acc_idx;      /* [npark]: byte set of a consuming instruction */
    std::vector<uint8_t>  kind;         /* [npark]: 0 consuming, 1 MATCH, 2 \z, 3 $, 4 \B, 5 \b */
    std::vector<uint16_t> regex;        /* [npark]: owning regex */
    std::vector<int32_t>  park_pc;      /* [npark]: its pc */
    std::vector<uint32_t> bent, bmask;  /* bucketed start closure (may be empty) */
    std::vector<uint16_t> bofs;         /* [3][257] */
    uint32_t              npark = 0;
    uint32_t              nsets = 0;
    uint32_t              max_slots = 0;        /* slots of the largest regex (<= 32) */
    int32_t               p_any = -1;           /* parked number of the ".*?" ANY (pc 1), -1: none */
    bool                  ctx_dep = false;      /* program has \A or ^ */
};

/* max_park: give up beyond that many parked instructions */
bool sre_build_closure_table(const sre_program_t *prog, uint32_t max_park, sre_closure_table_t &T);

#endif
