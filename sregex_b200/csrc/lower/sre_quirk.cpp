/* sre_quirk.cpp -- see sre_quirk.h */
#include "sre_quirk.h"

#include <string.h>
#include <vector>

namespace {

struct replay_t {
    const sre_program_t   *prog;
    std::vector<uint32_t>  tags;
    uint32_t               tag = 0;
    std::vector<int32_t>   list;        /* pcs of the threads appended, in order */
    bool                   done = false;

    /* add_thread without captures, no assertion in the program (:756-942) */
    void add(int32_t pc, bool want_done)
    {
        if (done) {
            return;
        }
        const sre_instruction_t &in = prog->insts[pc];
        if (tags[pc] == tag) {
            if (in.opcode == SRE_OPCODE_SPLIT && tags[in.y] != tag) {   /* the revisited-SPLIT rule */
                add(in.y, want_done);
            }
            return;
        }
        tags[pc] = tag;
        switch (in.opcode) {
        case SRE_OPCODE_JMP:
            add(in.x, want_done);
            return;
        case SRE_OPCODE_SPLIT:
            add(in.x, want_done);
            add(in.y, want_done);
            return;
        case SRE_OPCODE_SAVE:
            add(pc + 1, want_done);
            return;
        case SRE_OPCODE_MATCH:
            if (want_done) {
                done = true;
                return;
            }
            /* fall through */
        default:
            list.push_back(pc);
            return;
        }
    }
};

bool takes(const sre_program_t *prog, const sre_instruction_t &in, uint32_t b)
{
    bool inr = false;
    switch (in.opcode) {
    case SRE_OPCODE_CHAR: return b == in.ch;
    case SRE_OPCODE_ANY:  return true;
    case SRE_OPCODE_IN:
    case SRE_OPCODE_NOTIN:
        for (uint32_t j = 0; j < in.nranges; j++) {
            const sre_vm_range_t &r = prog->ranges[in.v + j];
            if (b >= r.from && b <= r.to) {
                inr = true;
                break;
            }
        }
        return inr == (in.opcode == SRE_OPCODE_IN);
    default: return false;
    }
}

}  // namespace

bool sre_quirk_bytes(const sre_program_t *prog, uint32_t single[8])
{
    memset(single, 0, 8 * sizeof(uint32_t));
    if (prog->nleading == 0) {
        return false;                   /* no prefilter (:256) */
    }
    uint32_t lead[8] = { 0 };
    for (uint32_t b = 0; b < 256; b++) {
        for (uint32_t i = 0; i < prog->nleading; i++) {
            if (takes(prog, prog->insts[prog->leading[i]], b)) {
                lead[b >> 5] |= 1u << (b & 31);
            }
        }
    }
    bool has_assert = false;
    for (uint32_t pc = 0; pc < prog->len; pc++) {
        has_assert |= prog->insts[pc].opcode == SRE_OPCODE_ASSERT;
    }
    if (has_assert) {
        /* closures depend on more than the byte: every leading byte is a candidate */
        memcpy(single, lead, sizeof(lead));
        return true;
    }
    replay_t R;
    R.prog = prog;
    R.tags.assign(prog->len + 1, 0);
    R.tag = 1;
    R.add(0, false);
    const std::vector<int32_t> init = R.list;
    bool any = false;
    for (uint32_t b = 0; b < 256; b++) {
        if (!((lead[b >> 5] >> (b & 31)) & 1)) {
            continue;                   /* a jump only lands on a leading byte */
        }
        R.tag++;
        R.list.clear();
        R.done = false;
        for (size_t i = 0; i < init.size() && !R.done; i++) {
            if (takes(prog, prog->insts[init[i]], b)) {
                R.add(init[i] + 1, true);
            }
        }
        if (!R.done || R.list.size() != init.size()) {
            continue;
        }
        bool same = true;
        for (size_t i = 0; i + 1 < init.size(); i++) {          /* every pc but the last, :262-274 */
            same &= R.list[i] == init[i];
        }
        if (same) {
            single[b >> 5] |= 1u << (b & 31);
            any = true;
        }
    }
    return any;
}
