/* sre_quirk.cpp -- see sre_quirk.h */
#include "sre_quirk.h"

#include <string.h>
#include <vector>

namespace {

/* what lies in front of the position a closure is computed at: the start of the input, or a byte */
const int PREV_NONE = -1;

struct qthread_t {
    int32_t pc;
    bool    seen_word;      /* of a thread parked on \b / \B (:868-887) */
};

/*
 * add_thread without captures (:756-942), the reference's one tag word per instruction
 * included: look-behind assertions (\A ^) are decided by `prev`, look-ahead ones ($ \z \b \B)
 * park the thread -- \b \B with the kind of `prev`.
 */
struct replay_t {
    const sre_program_t    *prog;
    std::vector<uint32_t>   tags;
    uint32_t                tag = 0;
    bool                    done = false;

    void add(std::vector<qthread_t> &list, int32_t pc, bool want_done, int prev)
    {
        if (done) {
            return;
        }
        const sre_instruction_t &in = prog->insts[pc];
        if (tags[pc] == tag) {
            if (in.opcode == SRE_OPCODE_SPLIT && tags[in.y] != tag) {   /* the revisited-SPLIT rule */
                add(list, in.y, want_done, prev);
            }
            return;
        }
        tags[pc] = tag;
        bool seen_word = false;
        switch (in.opcode) {
        case SRE_OPCODE_JMP:
            add(list, in.x, want_done, prev);
            return;
        case SRE_OPCODE_SPLIT:
            add(list, in.x, want_done, prev);
            add(list, in.y, want_done, prev);
            return;
        case SRE_OPCODE_SAVE:
            add(list, pc + 1, want_done, prev);
            return;
        case SRE_OPCODE_ASSERT:
            switch (in.v) {
            case SRE_REGEX_ASSERT_BIG_A:
                if (prev == PREV_NONE) {
                    add(list, pc + 1, want_done, prev);
                }
                return;
            case SRE_REGEX_ASSERT_CARET:
                if (prev == PREV_NONE || prev == '\n') {
                    add(list, pc + 1, want_done, prev);
                }
                return;
            case SRE_REGEX_ASSERT_SMALL_B:
            case SRE_REGEX_ASSERT_BIG_B:
                seen_word = prev != PREV_NONE && sre_isword(prev);
                break;
            default:
                break;              /* $ \z: postponed */
            }
            break;
        case SRE_OPCODE_MATCH:
            if (want_done) {
                done = true;
                return;
            }
            break;
        default:
            break;
        }
        list.push_back(qthread_t{ pc, seen_word });
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
    /* the list sre_vm_pike_exec records as "initial": the closure of the start at offset 0 (:202-229) */
    replay_t R;
    R.prog = prog;
    R.tags.assign(prog->len + 1, 0);
    R.tag = 1;
    std::vector<qthread_t> init;
    R.add(init, 0, false, PREV_NONE);
    bool has_assert = false;
    for (uint32_t pc = 0; pc < prog->len; pc++) {
        has_assert |= prog->insts[pc].opcode == SRE_OPCODE_ASSERT;
    }
    /* a jump lands at an offset >= 1: the byte in front of it is a newline, a word byte or neither
     * (one case when no assertion looks at it) */
    static const int prevs[3] = { ' ', 'a', '\n' };
    bool any = false;
    std::vector<qthread_t> cur, next, held;
    for (int pi = 0; pi < (has_assert ? 3 : 1); pi++) {
        const int prev = prevs[pi];
        for (uint32_t b = 0; b < 256; b++) {
            if (!((lead[b >> 5] >> (b & 31)) & 1) || ((single[b >> 5] >> (b & 31)) & 1)) {
                continue;               /* a jump only lands on a leading byte */
            }
            /* the fresh list the jump builds at the landing offset (:283-297) ... */
            R.tag += 2;
            R.done = false;
            cur.clear();
            R.add(cur, 0, false, prev);
            /* ... and the step on the landing byte (:309-560) */
            const uint32_t T = ++R.tag;
            next.clear();
            bool matched = false;
            for (size_t i = 0; i < cur.size() && !matched; i++) {
                const qthread_t t = cur[i];
                const sre_instruction_t &in = prog->insts[t.pc];
                if (in.opcode == SRE_OPCODE_MATCH) {
                    matched = true;                             /* a parked MATCH thread (:524-553) */
                } else if (in.opcode == SRE_OPCODE_ASSERT) {
                    bool hold = false;
                    switch (in.v) {
                    case SRE_REGEX_ASSERT_DOLLAR:  hold = b == '\n'; break;
                    case SRE_REGEX_ASSERT_SMALL_B: hold = t.seen_word != (bool) sre_isword(b); break;
                    case SRE_REGEX_ASSERT_BIG_B:   hold = t.seen_word == (bool) sre_isword(b); break;
                    default: break;                             /* \z: not in front of a byte */
                    }
                    if (hold) {
                        /* its closure at this offset, under the tag the current list was built with,
                         * goes in front of the threads still to run (:484-509) */
                        R.tag = T - 1;
                        held.clear();
                        R.add(held, t.pc + 1, false, prev);
                        R.tag = T;
                        cur.insert(cur.begin() + (long) i + 1, held.begin(), held.end());
                    }
                } else if (takes(prog, in, b)) {
                    R.add(next, t.pc + 1, true, (int) b);
                    matched = R.done;
                }
            }
            if (!matched || next.size() != init.size()) {
                continue;
            }
            bool same = true;
            for (size_t i = 0; i + 1 < init.size(); i++) {      /* every pc but the last, :262-274 */
                same &= next[i].pc == init[i].pc;
            }
            if (same) {
                single[b >> 5] |= 1u << (b & 31);
                any = true;
            }
        }
    }
    return any;
}
