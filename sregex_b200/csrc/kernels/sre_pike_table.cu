/*
 * sre_pike_table.cu -- the Pike VM driven by precomputed closure tables (the
 * batch Pike+captures kernel of choice: single regexes and regex sets whose
 * tables fit in shared memory).
 *
 * Same results as sre_vm_pike_exec (reference sre_vm_pike.c:148-689) for a
 * fresh context and one buffer with eof = 1, like sre_pike_small.cu, but
 * add_thread (:756-942) is not walked at run time.  Instructions a thread can
 * be parked on (consuming ones, look-ahead assertions, MATCH) are numbered
 * 0 .. npark-1 (SPLIT / JMP / SAVE and `\A` `^` never hold a thread), and
 * everything below speaks of those numbers.  For every such instruction P, the
 * host lists what add_thread(pc(P) + 1) appends when run on its own
 * (lower/sre_closure.cpp -- same walk order, same revisited-SPLIT rule
 * :770-786): the parked instructions in priority order, each with the set of
 * capture slots SAVEd on the way, once per look-behind context (at offset 0 /
 * after a newline / elsewhere, which is all `\A` and `^` can see).  At run time
 * a closure is "for each entry: skip it if its instruction is already marked in
 * this step, else mark it and append a thread whose captures are the parent's
 * with the SAVEd slots set to the position".  That is the list the walk would
 * have produced: a walk stops at an instruction marked earlier in the step, and
 * everything below such an instruction was appended (and marked) by the walk
 * that marked it.  The idea is the reference JIT's
 * (sre_vm_thompson_x64.dasc:323-394 precomputes the closure of every consuming
 * instruction), extended with captures.  oracle/lower_check.cpp holds a CPU
 * model of this kernel over the same tables (tests/test_lowering.py).
 *
 * On top of that:
 *   - a thread parked on a consuming instruction that cannot take the next byte
 *     is not appended (the next step would drop it without effect);
 *   - the ".*?" thread is a flag, not a list entry (it is always last, always
 *     takes the byte and carries no capture); its closure -- the start closure
 *     -- comes bucketed by the next byte when it is long (regex sets);
 *   - dedup marks are "this step / the previous step", the only two epochs the
 *     reference compares against: two 64-bit registers, or two bit sets in
 *     shared memory for programs with more than 64 parked instructions;
 *   - a thread carries the capture slots of its own regex only (<= 32), as
 *     16-bit line offsets when lines are shorter than 32 KB;
 *   - the first-byte prefilter is left to the start hint of k_dfa_lines_hint;
 *   - lines are handed out from a global counter, and every turn of the one
 *     (line, byte) loop starts with a warp vote that lines the lanes up again.
 *
 * A line that needs more than K threads per list or H pending look-ahead
 * closures is reported SRE_K_RETRY and re-run with larger lists, then by
 * k_pike_lines.
 */
#include "sre_kernels.cuh"

namespace {

constexpr int TB = 128;     /* threads (= contexts) per block */
enum { NB_END = -2 };

__device__ __forceinline__ bool isword(uint32_t c)
{
    return (c - '0' < 10u) || ((c | 0x20) - 'a' < 26u) || c == '_';
}

/*
 * one lane's context: words at sm[e * TB].  BIG: more than 64 parked
 * instructions -- the dedup marks are two bit sets in shared memory (MK, 2 x mw
 * words, the current one chosen by `par`) instead of two 64-bit registers.
 */
template <bool C16, bool BIG, int NCW>
struct lane_t {
    int32_t    *sm;
    int         ncw_rt;             /* words per capture vector (NCW == 0: only known at run time) */
    __device__ __forceinline__ int ncw_() const { return NCW ? NCW : ncw_rt; }
    int         K, H;               /* threads per list, pending look-ahead closures */
    uint64_t    m_cur, m_prev;
    int         mw, par;            /* BIG: words per mark set, parity of the current set */
    /* sections (word offsets), set from the vector width */
    int         MAT, TMP, L0PC, L0CAP, L1PC, L1CAP, HSPC, HSCAP, MK;

    __device__ __forceinline__ void layout(int n, int k, int h, int npark)
    {
        ncw_rt = n;
        K = k;
        H = h;
        mw = BIG ? (npark + 31) >> 5 : 0;
        par = 0;
        MAT = 0;
        TMP = MAT + n;
        L0PC = TMP + n;
        L0CAP = L0PC + K;
        L1PC = L0CAP + K * n;
        L1CAP = L1PC + K;
        HSPC = L1CAP + K * n;
        HSCAP = HSPC + H;
        MK = HSCAP + H * n;
    }
    __device__ __forceinline__ int32_t &w(int e) { return sm[e * TB]; }

    /* all marks off (new line) */
    __device__ __forceinline__ void marks_reset()
    {
        if (BIG) {
            for (int j = 0; j < 2 * mw; j++) {
                w(MK + j) = 0;
            }
            par = 0;
        } else {
            m_cur = m_prev = 0;
        }
    }
    /* ctx->tag++: the current set becomes the previous one, the new current one is empty */
    __device__ __forceinline__ void marks_advance()
    {
        if (BIG) {
            par ^= 1;
            for (int j = 0; j < mw; j++) {
                w(MK + par * mw + j) = 0;
            }
        } else {
            m_prev = m_cur;
            m_cur = 0;
        }
    }
    /* marked with ctx->tag (hold == false) or with ctx->tag - 1 */
    __device__ __forceinline__ bool tagged(uint32_t pc, bool hold)
    {
        if (BIG) {
            return ((uint32_t) w(MK + (par ^ (int) hold) * mw + (int) (pc >> 5)) >> (pc & 31)) & 1;
        }
        return ((hold ? m_prev : m_cur) >> pc) & 1;
    }
    __device__ __forceinline__ void tag(uint32_t pc, bool hold)
    {
        if (BIG) {
            const int32_t bit = (int32_t) (1u << (pc & 31));
            w(MK + (par ^ (int) hold) * mw + (int) (pc >> 5)) |= bit;
            w(MK + (par ^ (int) hold ^ 1) * mw + (int) (pc >> 5)) &= ~bit;
            return;
        }
        const uint64_t bit = 1ull << pc;
        if (hold) {
            m_prev |= bit;
            m_cur &= ~bit;
        } else {
            m_cur |= bit;
            m_prev &= ~bit;
        }
    }
    /* dst vector <- src vector (src < 0: all -1) with the slots of `mask` set to pos:
     * a plain copy, then one (half-)word store per SAVEd slot -- the same path
     * for every lane, whatever its mask */
    __device__ __forceinline__ void derive(int dst, int src, uint32_t mask, int32_t pos)
    {
        for (int j = 0; j < ncw_(); j++) {
            w(dst + j) = src < 0 ? -1 : w(src + j);
        }
        while (mask) {
            const uint32_t slot = (uint32_t) __ffs((int) mask) - 1;
            mask &= mask - 1;
            if (C16) {
                reinterpret_cast<int16_t *>(&w(dst + (int) (slot >> 1)))[slot & 1] = (int16_t) pos;
            } else {
                w(dst + (int) slot) = pos;
            }
        }
    }
    __device__ __forceinline__ int32_t cap_get(int sec, uint32_t slot)
    {
        if (!C16) {
            return w(sec + slot);
        }
        const int32_t word = w(sec + (slot >> 1));
        return (slot & 1) ? (word >> 16) : (int32_t) (int16_t) (word & 0xffff);
    }
};

/* what a parked instruction is (block table s_kind) */
enum { KD_CONS = 0, KD_MATCH = 1, KD_SMALL_Z = 2, KD_DOLLAR = 3, KD_BIG_B = 4, KD_SMALL_B = 5 };

/* block tables in shared memory, in this order (words, then halves, then bytes) */
struct tables_t {
    uint32_t  *ent, *emask, *bent, *bmask, *accept;
    uint16_t  *ofs, *bofs, *accidx, *regex;
    uint8_t   *kind;
};

__host__ __device__ inline size_t table_words(const sre_dev_pike_t &pk)
{
    const size_t np = pk.clo_npark;
    const size_t halves = 3 * (np + 2) + (pk.clo_nbent ? 3 * 257 : 0) + 2 * np;
    return 2 * (size_t) (pk.clo_nent + pk.clo_nbent) + (size_t) pk.clo_nsets * 8 + (halves * 2 + np + 3) / 4;
}

/* HOLD: the program has look-ahead assertions ($ \z \b \B), whose closures are
 * prepended to the current list through the LIFO; without them that code is
 * compiled out of the thread loop */
template <bool C16, bool BIG, bool HOLD, int NCW>
__global__ void __launch_bounds__(TB)
k_pike_table(sre_dev_pike_t pk, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
             size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines,
             const int32_t *__restrict__ start_hint, int32_t *__restrict__ rc, int64_t *__restrict__ ovec,
             uint32_t ovec_slots, int K, int H, int pass, sre_pike_work_t *work)
{
    extern __shared__ int32_t smem_words[];
    /* the retry pass has nothing to do when the first pass gave up on no line */
    const bool retry_only = pass != 0;
    if (retry_only && work->given_up[0] == 0) {
        return;
    }
    unsigned long long *next_work = &work->next;
    const uint32_t len = pk.clo_npark, nofs = 3 * (len + 2), nbofs = pk.clo_nbent ? 3 * 257 : 0;
    tables_t t;
    t.ent = reinterpret_cast<uint32_t *>(smem_words);
    t.emask = t.ent + pk.clo_nent;
    t.bent = t.emask + pk.clo_nent;
    t.bmask = t.bent + pk.clo_nbent;
    t.accept = t.bmask + pk.clo_nbent;
    t.ofs = reinterpret_cast<uint16_t *>(t.accept + pk.clo_nsets * 8);
    t.bofs = t.ofs + nofs;
    t.accidx = t.bofs + nbofs;
    t.regex = t.accidx + len;
    t.kind = reinterpret_cast<uint8_t *>(t.regex + len);
    for (uint32_t i = threadIdx.x; i < pk.clo_nent; i += TB) {
        t.ent[i] = pk.clo_ent[i];
        t.emask[i] = pk.clo_emask[i];
    }
    for (uint32_t i = threadIdx.x; i < pk.clo_nbent; i += TB) {
        t.bent[i] = pk.clo_bent[i];
        t.bmask[i] = pk.clo_bmask[i];
    }
    for (uint32_t i = threadIdx.x; i < pk.clo_nsets * 8; i += TB) {
        t.accept[i] = pk.clo_accept[i];
    }
    for (uint32_t i = threadIdx.x; i < nofs; i += TB) {
        t.ofs[i] = pk.clo_ofs[i];
    }
    for (uint32_t i = threadIdx.x; i < nbofs; i += TB) {
        t.bofs[i] = pk.clo_bofs[i];
    }
    for (uint32_t i = threadIdx.x; i < len; i += TB) {
        t.accidx[i] = pk.clo_accidx[i];
        t.regex[i] = pk.clo_regex[i];
        t.kind[i] = pk.clo_kind[i];
    }
    __syncthreads();
    const uint32_t *s_ent = t.ent, *s_emask = t.emask, *s_bent = t.bent, *s_bmask = t.bmask, *s_accept = t.accept;
    const uint16_t *s_ofs = t.ofs, *s_bofs = t.bofs, *s_accidx = t.accidx, *s_regex = t.regex;
    const uint8_t *s_kind = t.kind;
    const uint32_t p_any = pk.clo_p_any;

    lane_t<C16, BIG, NCW> c;
    c.sm = smem_words + table_words(pk) + threadIdx.x;
    /* a thread carries the slots of its own regex only */
    c.layout(C16 ? (int) (pk.max_slots + 1) >> 1 : (int) pk.max_slots, K, H, (int) len);
    const int ncw = c.ncw_();
    const bool ctx_dep = pk.clo_ctx_dep != 0;

    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    const size_t nwork = lines.list ? (size_t) *lines.count : nlines;
    size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x;

    /* the line this lane is on */
    size_t line = 0;
    const uint8_t *input = buf;
    int32_t size = 0, sp = 0, matched_id = 0;
    bool active = false, overflow = false, matched = false;
    /* the ".*?" thread is not kept in the lists: it is always the last thread of
     * a list, always takes the byte and never carries a capture, so one flag
     * stands for it (off once a match has cut the lower-priority threads) */
    bool any_alive = false;
    int cur = 0, ncl = 0, nnl = 0, hs = 0;
    /* the byte at sp (NB_END at the end of the line) and the one before it,
     * carried from step to step: one load per step */
    int nb_cur = NB_END;
    uint32_t prev_byte = 0;

    /*
     * closure P appended to the array (pcsec, capsec) of capacity capn at
     * count n: 0 ok, 1 MATCH reached (want_done), -1 out of room
     */
    /* prev / nb: the bytes before and at `pos` (nb: NB_END at the end of the line);
     * the caller has them at hand for the whole step */
    auto append_closure = [&](uint32_t P, int32_t pos, uint32_t prev, int nb, int parent, int pcsec, int capsec,
                              int capn, int &n, bool hold, bool want_done) -> int {
        uint32_t v = 0;
        if (pos > 0) {
            v = ctx_dep ? (prev == '\n' ? 1u : 2u) : 0u;
        }
        const uint32_t *list = s_ent, *lmask = s_emask;
        uint32_t e = s_ofs[v * (len + 2) + P], e1 = s_ofs[v * (len + 2) + P + 1];
        bool filtered = false;
        if ((P == len || P == p_any) && pk.clo_nbent && nb >= 0) {
            /* the start closure, already restricted to this next byte */
            list = s_bent;
            lmask = s_bmask;
            e = s_bofs[v * 257 + (uint32_t) nb];
            e1 = s_bofs[v * 257 + (uint32_t) nb + 1];
            filtered = true;
        }
        for (; e < e1; e++) {
            const uint32_t fpc = list[e];
            if (fpc == p_any) {
                continue;               /* the ".*?" thread itself: see any_alive */
            }
            const uint32_t kind = s_kind[fpc];
            if (!filtered && kind == KD_CONS
                && (nb == NB_END
                    || !((s_accept[(uint32_t) s_accidx[fpc] * 8 + ((uint32_t) nb >> 5)] >> (nb & 31)) & 1)))
            {
                continue;               /* would be dropped by the next step */
            }
            if (c.tagged(fpc, hold)) {
                continue;
            }
            c.tag(fpc, hold);
            const uint32_t mask = lmask[e];
            if (kind == KD_MATCH && want_done) {
                c.derive(c.MAT, parent, mask, pos);
                matched_id = (int32_t) s_regex[fpc];
                return 1;
            }
            if (n >= capn) {
                return -1;
            }
            const uint32_t sw = (kind >= KD_BIG_B && pos > 0 && isword(prev)) ? 1u : 0u;
            c.w(pcsec + n) = (int32_t) (fpc | (sw << 16));
            c.derive(capsec + n * ncw, parent, mask, pos);
            n++;
        }
        return 0;
    };

    /*
     * One loop over (line, byte) pairs: a lane that finishes its line takes
     * the next one at once instead of waiting for the longest line of the
     * warp, so every turn is one byte step for every lane that has work left.
     */
    /*
     * Work is handed out line by line from a global counter (the first line of
     * every lane is its own index): a lane that finishes early takes the next
     * line of the whole launch, so short and long lines even out whatever the
     * number of lines per lane.  One atomic per warp and turn, for all lanes
     * that need a line.
     */
    bool finished = false;
    bool first_line = true;
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        /* the votes are also the point where the lanes of the warp line up
         * again: without it they drift apart and run one after the other */
        if (__all_sync(0xffffffffu, finished)) {
            break;
        }
        const bool need = !finished && !active && !first_line;
        const uint32_t needers = __ballot_sync(0xffffffffu, need);
        if (needers) {
            unsigned long long base = 0;
            if (lane == (uint32_t) (__ffs(needers) - 1)) {
                base = atomicAdd(next_work, (unsigned long long) __popc(needers));
            }
            base = __shfl_sync(0xffffffffu, base, __ffs(needers) - 1);
            if (need) {
                k = nthreads + (size_t) base + __popc(needers & ((1u << lane) - 1));
            }
        }
        if (finished) {
            continue;
        }
        if (!active) {
            first_line = false;
            if (k >= nwork) {
                finished = true;
                continue;
            }
            line = lines.list ? (size_t) lines.list[k] : k;
            /* a later pass with larger lists: only what the previous one gave up on */
            if (retry_only && rc[line] != SRE_K_RETRY) {
                continue;
            }
            const size_t start = offsets ? (size_t) offsets[line] : line * pitch;
            const size_t end = offsets ? (size_t) offsets[line + 1] : start + linelen;
            input = buf + start;
            size = (int32_t) (end - start);
            sp = start_hint ? start_hint[line] : 0;
            c.marks_reset();
            overflow = false;
            matched = false;
            matched_id = 0;
            cur = 0; ncl = 0; nnl = 0; hs = 0;
            any_alive = p_any != 0xffffffffu;
            active = true;
            if (C16 && size > 32766) {      /* longer than the caller's bound: not with 16-bit offsets */
                overflow = true;
            }
            /* first_buf: the initial closure at the start offset, :202-216 */
            prev_byte = sp > 0 ? (uint32_t) input[sp - 1] : 0u;
            nb_cur = sp < size ? (int) input[sp] : NB_END;
            if (!overflow
                && append_closure(len, sp, prev_byte, nb_cur, -1, c.L0PC, c.L0CAP, c.K, ncl, false, false) < 0)
            {
                overflow = true;
            }
        }

        bool done = overflow || sp > size || (ncl == 0 && !any_alive);
        if (!done) {
            c.marks_advance();          /* ctx->tag++ */
            const bool at_end = (sp == size);
            const uint32_t byte = at_end ? 0 : (uint32_t) nb_cur;
            const bool cur_word = !at_end && isword(byte);
            /* the byte the threads appended in this step will be stepped on */
            const int nb_next = sp + 1 < size ? (int) input[sp + 1] : NB_END;
            const int cl_pc = cur ? c.L1PC : c.L0PC, cl_cap = cur ? c.L1CAP : c.L0CAP;
            const int nl_pc = cur ? c.L0PC : c.L1PC, nl_cap = cur ? c.L0CAP : c.L1CAP;
            int i = 0;

            for (;;) {
                /* next thread in priority order: pending look-ahead closures first */
                int tp, tc;
                if (HOLD && hs > 0) {
                    hs--;
                    tp = c.HSPC + hs;
                    /* its closure may be appended over this very record */
                    for (int j = 0; j < ncw; j++) {
                        c.w(c.TMP + j) = c.w(c.HSCAP + hs * ncw + j);
                    }
                    tc = c.TMP;
                } else if (i < ncl) {
                    tp = cl_pc + i;
                    tc = cl_cap + i * ncw;
                    i++;
                } else {
                    break;
                }
                const int32_t rec = c.w(tp);
                const uint32_t pc = rec & 0xffff;
                const bool t_sw = (rec >> 16) & 1;
                const uint32_t kind = s_kind[pc];
                bool got_match = false;

                if (HOLD && kind >= KD_SMALL_Z) {           /* :449-528 */
                    bool hold;
                    switch (kind) {
                    case KD_SMALL_Z: hold = at_end; break;
                    case KD_DOLLAR:  hold = at_end || byte == '\n'; break;
                    case KD_BIG_B:   hold = (t_sw == cur_word); break;
                    default:         hold = (t_sw != cur_word); break;
                    }
                    if (hold) {
                        /* closure with tag - 1, prepended to clist: append it
                         * above the LIFO top, then reverse that segment */
                        int top = hs;
                        if (append_closure(pc, sp, prev_byte, at_end ? NB_END : (int) byte, tc, c.HSPC, c.HSCAP,
                                           c.H, top, true, false) < 0)
                        {
                            overflow = true;
                            break;
                        }
                        for (int lo = hs, hi = top - 1; lo < hi; lo++, hi--) {
                            int32_t tmp = c.w(c.HSPC + lo);
                            c.w(c.HSPC + lo) = c.w(c.HSPC + hi);
                            c.w(c.HSPC + hi) = tmp;
                            for (int j = 0; j < ncw; j++) {
                                tmp = c.w(c.HSCAP + lo * ncw + j);
                                c.w(c.HSCAP + lo * ncw + j) = c.w(c.HSCAP + hi * ncw + j);
                                c.w(c.HSCAP + hi * ncw + j) = tmp;
                            }
                        }
                        hs = top;
                    }
                } else if (kind == KD_MATCH) {              /* :530-553 */
                    for (int j = 0; j < ncw; j++) {
                        c.w(c.MAT + j) = c.w(tc + j);
                    }
                    matched_id = (int32_t) s_regex[pc];
                    got_match = true;
                } else if (!at_end && ((s_accept[(uint32_t) s_accidx[pc] * 8 + (byte >> 5)] >> (byte & 31)) & 1)) {
                    const int r = append_closure(pc, sp + 1, byte, nb_next, tc, nl_pc, nl_cap, c.K, nnl, false, true);
                    if (r < 0) {
                        overflow = true;
                        break;
                    }
                    got_match = (r == 1);
                }

                if (got_match) {        /* cut every lower-priority thread */
                    matched = true;
                    any_alive = false;
                    hs = 0;
                    break;
                }
            }
            /* the ".*?" thread, last in priority: takes the byte and restarts the regex */
            if (any_alive && !at_end && !overflow) {
                const int r = append_closure(p_any, sp + 1, byte, nb_next, -1, nl_pc, nl_cap, c.K, nnl, false, true);
                if (r < 0) {
                    overflow = true;
                } else if (r == 1) {
                    matched = true;
                    any_alive = false;
                }
            }

            cur ^= 1;                   /* step_done: swap lists */
            ncl = nnl;
            nnl = 0;
            hs = 0;
            sp++;
            prev_byte = byte;
            nb_cur = nb_next;
            done = overflow || at_end || (ncl == 0 && !any_alive);
        }

        if (done) {
            int64_t *ov = ovec + line * ovec_slots;
            if (overflow) {
                rc[line] = SRE_K_RETRY;
                atomicAdd(&work->given_up[pass], 1u);
            } else if (matched) {
                /* prepare_matched_captures :945-989: the matched regex's slots, the rest -1 */
                const uint32_t cnt = pk.slot_ofs[matched_id + 1] - pk.slot_ofs[matched_id];
                rc[line] = matched_id;
                for (uint32_t i = 0; i < ovec_slots; i++) {
                    ov[i] = i < cnt ? (int64_t) c.cap_get(c.MAT, i) : -1;
                }
            } else {
                rc[line] = SRE_K_DECLINED;
                for (uint32_t i = 0; i < ovec_slots; i++) {
                    ov[i] = -1;
                }
            }
            active = false;
        }
    }
}

size_t table_smem_bytes(const sre_dev_pike_t &pk, bool c16, int K, int H)
{
    const bool big = pk.clo_npark > 64;
    const int ncw = c16 ? (int) (pk.max_slots + 1) >> 1 : (int) pk.max_slots;
    const size_t lane_words = 2 * ncw + 2 * K * (1 + ncw) + H * (1 + ncw)
                              + (big ? 2 * ((pk.clo_npark + 31) >> 5) : 0);
    return (table_words(pk) + lane_words * TB) * 4;
}

}  // namespace

/* 16-bit capture offsets when every line is shorter than 32 KB: linelen is the
 * line length (fixed pitch) or, with offsets, the caller's upper bound on it
 * (0 = none given).  A longer line met anyway is handed to the next tier. */
static bool use_c16(const int64_t *offsets, size_t linelen)
{
    (void) offsets;
    return linelen != 0 && linelen < 32767;
}

bool sre_pike_table_applicable(const sre_dev_pike_t &pk, const int64_t *offsets, size_t linelen, int K, int H)
{
    return pk.clo_nent != 0 && pk.max_slots <= 32
           && table_smem_bytes(pk, use_c16(offsets, linelen), K, H) <= 200 * 1024;
}

cudaError_t sre_launch_pike_table(const sre_dev_pike_t &pk, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines, const int32_t *start, int32_t *rc,
    int64_t *ovec, uint32_t ovec_slots, int K, int H, int pass, sre_pike_work_t *work,
    cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    /* first pass: everything; retry pass: the work counter only */
    cudaError_t ce = cudaMemsetAsync(work, 0, pass == 0 ? sizeof(*work) : sizeof(work->next), stream);
    if (ce != cudaSuccess) {
        return ce;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
            sms = 148;
        }
    }
    const bool c16 = use_c16(offsets, linelen);
    const bool big = pk.clo_npark > 64;
    const size_t smem = table_smem_bytes(pk, c16, K, H);
    size_t per_sm = (227 * 1024) / (smem + 1024);
    if (per_sm > 16) {
        per_sm = 16;
    }
    size_t grid = (nlines + TB - 1) / TB;
    const size_t cap = (size_t) sms * (per_sm ? per_sm : 1);
    if (grid > cap) {
        grid = cap;
    }
    typedef void (*kern_t)(sre_dev_pike_t, const uint8_t *, const int64_t *, size_t, size_t, size_t,
                           sre_line_list_t, const int32_t *, int32_t *, int64_t *, uint32_t, int, int, int,
                           sre_pike_work_t *);
    /* [c16][big][hold][capture words: 0 = run-time, 1, 3, 5 (16-bit) / 2, 6, 10 (32-bit)] */
#define SRE_TAB_ROW(C, B, H, N1, N2, N3)                                                            \
    { k_pike_table<C, B, H, 0>, k_pike_table<C, B, H, N1>, k_pike_table<C, B, H, N2>, k_pike_table<C, B, H, N3> }
    static const kern_t kerns[8][4] = {
        SRE_TAB_ROW(false, false, false, 2, 6, 10), SRE_TAB_ROW(false, false, true, 2, 6, 10),
        SRE_TAB_ROW(false, true, false, 2, 6, 10),  SRE_TAB_ROW(false, true, true, 2, 6, 10),
        SRE_TAB_ROW(true, false, false, 1, 3, 5),   SRE_TAB_ROW(true, false, true, 1, 3, 5),
        SRE_TAB_ROW(true, true, false, 1, 3, 5),    SRE_TAB_ROW(true, true, true, 1, 3, 5),
    };
#undef SRE_TAB_ROW
    const int which = (c16 ? 4 : 0) + (big ? 2 : 0) + (pk.clo_has_hold ? 1 : 0);
    /* 0, 2 or 4 capture groups get unrolled capture loops */
    const int sel = pk.max_slots == 2 ? 1 : pk.max_slots == 6 ? 2 : pk.max_slots == 10 ? 3 : 0;
    const kern_t kern = kerns[which][sel];
    static bool opted[8][4];
    if (!opted[which][sel]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            return e;
        }
        opted[which][sel] = true;
    }
    kern<<<(unsigned) grid, TB, smem, stream>>>(pk, buf, offsets, nlines, pitch, linelen, lines, start, rc, ovec,
                                               ovec_slots, K, H, pass, work);
    return cudaGetLastError();
}
