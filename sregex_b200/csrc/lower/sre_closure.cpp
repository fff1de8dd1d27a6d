/*
 * sre_closure.cpp -- see sre_closure.h.
 */
#include "sre_closure.h"

#include <string.h>

/*
 * The threads add_thread(pc 0) parks on consuming instructions, in the order
 * the reference's closure walk (sre_vm_pike.c:756-942: x before y, the
 * revisited-SPLIT rule :770-786) reaches them, each with the slots SAVEd on
 * its path; then bucketed by the byte they can take.  The Pike kernel appends
 * bucket[next byte] instead of walking the closure at every position.  Only
 * for programs whose start closure is context free: nleading != 0 (no MATCH,
 * no ANY) and no assertion on the way.
 */
bool sre_build_start_closure(const sre_program_t *prog, const std::vector<uint16_t> &pc_regex,
    const std::vector<uint32_t> &slot_ofs, std::vector<uint32_t> &ofs, std::vector<sre_start_ent_t> &ents)
{
    if (!prog->nleading || prog->len < 3 || prog->insts[0].opcode != SRE_OPCODE_SPLIT
        || prog->insts[0].y != 1 || prog->insts[1].opcode != SRE_OPCODE_ANY)
    {
        return false;
    }
    struct item_t { int32_t kind, pc; };            /* kind -1: visit pc; else undo one SAVE */
    std::vector<item_t> stack;
    std::vector<int32_t> saved;                     /* slots SAVEd on the current path */
    std::vector<uint8_t> seen(prog->len, 0);
    std::vector<sre_start_ent_t> finals;
    seen[0] = 1;
    stack.push_back({ -1, prog->insts[0].x });
    while (!stack.empty()) {
        const item_t it = stack.back();
        stack.pop_back();
        if (it.kind >= 0) {
            saved.pop_back();
            continue;
        }
        int32_t pc = it.pc;
        for (;;) {
            if (pc < 0 || (uint32_t) pc >= prog->len) {
                return false;
            }
            const sre_instruction_t &in = prog->insts[pc];
            if (seen[pc]) {
                if (in.opcode == SRE_OPCODE_SPLIT && !seen[in.y]) {
                    pc = in.y;
                    continue;
                }
                break;
            }
            seen[pc] = 1;
            if (in.opcode == SRE_OPCODE_JMP) {
                pc = in.x;
                continue;
            }
            if (in.opcode == SRE_OPCODE_SPLIT) {
                stack.push_back({ -1, in.y });
                pc = in.x;
                continue;
            }
            if (in.opcode == SRE_OPCODE_SAVE) {
                stack.push_back({ 0, 0 });
                saved.push_back(in.v);
                pc++;
                continue;
            }
            if (in.opcode == SRE_OPCODE_ASSERT || in.opcode == SRE_OPCODE_MATCH
                || in.opcode == SRE_OPCODE_ANY)
            {
                return false;
            }
            sre_start_ent_t e;
            memset(&e, 0, sizeof(e));
            e.pc = pc;
            const uint32_t base = slot_ofs[pc_regex[pc]], end = slot_ofs[pc_regex[pc] + 1];
            for (int32_t v : saved) {
                if ((uint32_t) v < base || (uint32_t) v >= end || (uint32_t) v - base > 255) {
                    return false;
                }
                const uint8_t rel = (uint8_t) ((uint32_t) v - base);
                bool dup = false;
                for (uint32_t k = 0; k < e.nsl; k++) {
                    dup |= (e.sl[k] == rel);
                }
                if (dup) {
                    continue;
                }
                if (e.nsl == sizeof(e.sl)) {
                    return false;
                }
                e.sl[e.nsl++] = rel;
            }
            finals.push_back(e);
            break;
        }
    }
    ofs.assign(257, 0);
    ents.clear();
    for (uint32_t b = 0; b < 256; b++) {
        ofs[b] = (uint32_t) ents.size();
        for (const sre_start_ent_t &e : finals) {
            const sre_instruction_t &in = prog->insts[e.pc];
            bool hit = false;
            if (in.opcode == SRE_OPCODE_CHAR) {
                hit = (in.ch == b);
            } else {
                for (uint32_t j = 0; j < in.nranges; j++) {
                    const sre_vm_range_t &r = prog->ranges[in.v + j];
                    hit |= (b >= r.from && b <= r.to);
                }
                if (in.opcode == SRE_OPCODE_NOTIN) {
                    hit = !hit;
                }
            }
            if (hit) {
                ents.push_back(e);
            }
        }
    }
    ofs[256] = (uint32_t) ents.size();
    return true;
}

/*
 * Closure tables (see sre_closure.h): the walk of sre_vm_pike.c:756-942 (x
 * before y, revisited-SPLIT rule :770-786, SAVE undone on the way back) run on
 * its own from every instruction a thread can be parked on.
 */

namespace {

struct walk_env_t {
    const sre_program_t          *prog;
    const std::vector<int32_t>   *park;
    const std::vector<uint16_t>  *pc_regex;
    const std::vector<uint32_t>  *slot_ofs;
};

/* false: a SAVE outside the 32-slot window of its regex */
bool closure_walk(const walk_env_t &env, int32_t pc0, int ctx, std::vector<uint32_t> &out,
    std::vector<uint32_t> &out_mask)
{
    const sre_program_t *prog = env.prog;
    struct item_t { int32_t kind, pc; uint32_t mask; };
    std::vector<item_t> stack;
    std::vector<uint8_t> seen(prog->len, 0);
    uint32_t mask = 0;
    stack.push_back({ -1, pc0, 0 });
    while (!stack.empty()) {
        const item_t it = stack.back();
        stack.pop_back();
        if (it.kind >= 0) {
            mask = it.mask;
            continue;
        }
        int32_t pc = it.pc;
        for (;;) {
            if (pc < 0 || (uint32_t) pc >= prog->len) {
                break;
            }
            const sre_instruction_t &in = prog->insts[pc];
            if (seen[pc]) {
                if (in.opcode == SRE_OPCODE_SPLIT && !seen[in.y]) {
                    pc = in.y;
                    continue;
                }
                break;
            }
            seen[pc] = 1;
            if (in.opcode == SRE_OPCODE_JMP) {
                pc = in.x;
                continue;
            }
            if (in.opcode == SRE_OPCODE_SPLIT) {
                stack.push_back({ -1, in.y, 0 });
                pc = in.x;
                continue;
            }
            if (in.opcode == SRE_OPCODE_SAVE) {
                const uint32_t rel = (uint32_t) in.v - (*env.slot_ofs)[(*env.pc_regex)[pc]];
                if (rel >= 32) {
                    return false;
                }
                stack.push_back({ 0, 0, mask });
                mask |= 1u << rel;
                pc++;
                continue;
            }
            if (in.opcode == SRE_OPCODE_ASSERT && in.v == SRE_REGEX_ASSERT_BIG_A) {
                if (ctx != 0) {
                    break;
                }
                pc++;
                continue;
            }
            if (in.opcode == SRE_OPCODE_ASSERT && in.v == SRE_REGEX_ASSERT_CARET) {
                if (ctx == 2) {
                    break;
                }
                pc++;
                continue;
            }
            out.push_back((uint32_t) (*env.park)[pc]);      /* parked */
            out_mask.push_back(mask);
            break;
        }
    }
    return true;
}

}  // namespace

bool sre_build_closure_table(const sre_program_t *prog, uint32_t max_park, sre_closure_table_t &T)
{
    const uint32_t len = prog->len;
    if (prog->nregexes == 0 || prog->nregexes > 65535) {
        return false;
    }
    /* the regex that owns a pc (code regions end at their MATCH) and its slots */
    std::vector<uint32_t> slot_ofs(prog->nregexes + 1, 0);
    for (sre_uint_t i = 0; i < prog->nregexes; i++) {
        const uint32_t cnt = 2 * (uint32_t) (prog->multi_ncaps[i] + 1);
        if (cnt > 32) {
            return false;
        }
        if (cnt > T.max_slots) {
            T.max_slots = cnt;
        }
        slot_ofs[i + 1] = slot_ofs[i] + cnt;
    }
    std::vector<uint16_t> pc_regex(len + 1, 0);
    {
        uint32_t r = 0;
        for (uint32_t pc = 0; pc < len; pc++) {
            pc_regex[pc] = (uint16_t) (r < prog->nregexes ? r : prog->nregexes - 1);
            if (prog->insts[pc].opcode == SRE_OPCODE_MATCH) {
                r++;
            }
        }
    }
    /* number the instructions that can hold a thread */
    std::vector<int32_t> park(len, -1);
    std::vector<int32_t> &park_pc = T.park_pc;
    park_pc.clear();
    for (uint32_t pc = 0; pc < len; pc++) {
        const sre_instruction_t &in = prog->insts[pc];
        uint8_t kind = 0xff;
        switch (in.opcode) {
        case SRE_OPCODE_CHAR:
        case SRE_OPCODE_ANY:
        case SRE_OPCODE_IN:
        case SRE_OPCODE_NOTIN:
            kind = 0;
            break;
        case SRE_OPCODE_MATCH:
            kind = 1;
            break;
        case SRE_OPCODE_ASSERT:
            switch (in.v) {
            case SRE_REGEX_ASSERT_SMALL_Z: kind = 2; break;
            case SRE_REGEX_ASSERT_DOLLAR:  kind = 3; break;
            case SRE_REGEX_ASSERT_BIG_B:   kind = 4; break;
            case SRE_REGEX_ASSERT_SMALL_B: kind = 5; break;
            default: T.ctx_dep = true; break;
            }
            break;
        default:
            break;
        }
        if (kind != 0xff) {
            park[pc] = (int32_t) park_pc.size();
            park_pc.push_back((int32_t) pc);
            T.kind.push_back(kind);
            T.regex.push_back(in.opcode == SRE_OPCODE_MATCH ? (uint16_t) in.v : pc_regex[pc]);
        }
    }
    const uint32_t np = (uint32_t) park_pc.size();
    if (np == 0 || np > max_park || np > 60000) {
        return false;
    }
    T.npark = np;
    if (len >= 3 && prog->insts[0].opcode == SRE_OPCODE_SPLIT && prog->insts[0].y == 1
        && prog->insts[1].opcode == SRE_OPCODE_ANY)
    {
        T.p_any = park[1];
    }
    /* byte sets of the consuming instructions, identical ones shared */
    T.acc_idx.assign(np, 0);
    for (uint32_t P = 0; P < np; P++) {
        const sre_instruction_t &in = prog->insts[park_pc[P]];
        if (T.kind[P] != 0) {
            continue;
        }
        uint32_t set[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (uint32_t b = 0; b < 256; b++) {
            bool hit;
            if (in.opcode == SRE_OPCODE_CHAR) {
                hit = (in.ch == b);
            } else if (in.opcode == SRE_OPCODE_ANY) {
                hit = true;
            } else {
                hit = false;
                for (uint32_t j = 0; j < in.nranges; j++) {
                    const sre_vm_range_t &r = prog->ranges[in.v + j];
                    hit |= (b >= r.from && b <= r.to);
                }
                if (in.opcode == SRE_OPCODE_NOTIN) {
                    hit = !hit;
                }
            }
            if (hit) {
                set[b >> 5] |= 1u << (b & 31);
            }
        }
        uint32_t k = 0;
        for (; k < T.nsets; k++) {
            if (memcmp(&T.accept[(size_t) k * 8], set, sizeof(set)) == 0) {
                break;
            }
        }
        if (k == T.nsets) {
            T.accept.insert(T.accept.end(), set, set + 8);
            T.nsets++;
        }
        T.acc_idx[P] = (uint16_t) k;
    }
    if (T.nsets == 0) {
        T.accept.assign(8, 0);
        T.nsets = 1;
    }

    const walk_env_t env = { prog, &park, &pc_regex, &slot_ofs };
    T.ofs.assign((size_t) 3 * (np + 2), 0);
    const int nctx = T.ctx_dep ? 3 : 1;
    for (int ctx = 0; ctx < nctx; ctx++) {
        for (uint32_t P = 0; P <= np; P++) {
            T.ofs[(size_t) ctx * (np + 2) + P] = (uint16_t) T.ent.size();
            bool ok = true;
            if (P == np) {
                ok = closure_walk(env, 0, ctx, T.ent, T.emask);
            } else if (T.kind[P] != 1) {            /* nothing follows a MATCH */
                ok = closure_walk(env, park_pc[P] + 1, ctx, T.ent, T.emask);
            }
            if (!ok || T.ent.size() > 60000) {
                return false;
            }
        }
        T.ofs[(size_t) ctx * (np + 2) + np + 1] = (uint16_t) T.ent.size();
    }
    for (int ctx = nctx; ctx < 3; ctx++) {
        for (uint32_t P = 0; P <= np + 1; P++) {
            T.ofs[(size_t) ctx * (np + 2) + P] = T.ofs[P];
        }
    }
    if (T.ent.empty()) {
        return false;
    }

    /* the start closure by next byte, when it is long enough to matter */
    T.bofs.assign((size_t) 3 * 257, 0);
    const uint32_t s0 = T.ofs[np], s1 = T.ofs[np + 1];
    if (s1 - s0 > 8) {
        for (int ctx = 0; ctx < nctx; ctx++) {
            const uint32_t e0 = T.ofs[(size_t) ctx * (np + 2) + np], e1 = T.ofs[(size_t) ctx * (np + 2) + np + 1];
            for (uint32_t b = 0; b < 256; b++) {
                T.bofs[(size_t) ctx * 257 + b] = (uint16_t) T.bent.size();
                for (uint32_t e = e0; e < e1; e++) {
                    const uint32_t P = T.ent[e];
                    if (T.kind[P] == 0
                        && !((T.accept[(size_t) T.acc_idx[P] * 8 + (b >> 5)] >> (b & 31)) & 1))
                    {
                        continue;
                    }
                    T.bent.push_back(T.ent[e]);
                    T.bmask.push_back(T.emask[e]);
                }
                if (T.bent.size() > 4096) {
                    /* buckets that are not selective would only cost shared
                     * memory: the tables are fine without them */
                    T.bent.clear();
                    T.bmask.clear();
                    return true;
                }
            }
            T.bofs[(size_t) ctx * 257 + 256] = (uint16_t) T.bent.size();
        }
        for (int ctx = nctx; ctx < 3; ctx++) {
            for (uint32_t b = 0; b <= 256; b++) {
                T.bofs[(size_t) ctx * 257 + b] = T.bofs[b];
            }
        }
    }
    return true;
}
