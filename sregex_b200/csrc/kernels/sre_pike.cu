/*
 * sre_pike.cu -- Pike VM (leftmost-first match, sub-match captures, matched
 * regex id) for sm_100a: one CUDA thread runs one context.
 *
 * What it replaces: sre_vm_pike_exec and helpers (reference sre_vm_pike.c:
 * exec :148-689, add_thread :756-942, prepare_matched_captures :945-989,
 * prepare_temp_captures :692-735, find_first_byte :992-1061) and the
 * copy-on-write capture vectors of sre_capture.c.
 *
 * The Perl-priority result of the reference is a property of the exact order
 * in which it walks its thread lists (DFS pre-order of the closure, the
 * revisited-SPLIT rule :770-786, "cut lower priority on match" :535-553,
 * look-ahead closures prepended with the current list's tag :506-525), so the
 * kernel keeps that sequential algorithm per context and is parallel across
 * contexts (= lines).  Differences in mechanism only:
 *   - the program is the read-only flat bytecode; the dedup marks are two bit
 *     sets per context selected by the parity of the tag counter (only the
 *     current and the previous epoch are ever consulted), in shared memory in
 *     the batch kernels;
 *   - add_thread's recursion is an explicit DFS stack with an undo log for the
 *     SAVE slots, which gives each branch of a SPLIT the captures the reference
 *     gives it through copy-on-write;
 *   - every list thread owns a private copy of the capture slots of the regex
 *     its pc belongs to (a thread inside regex i of a multi-regex set can only
 *     have written regex i's slots; all others are -1 by construction), so a
 *     thread costs max_slots instead of nslots values;
 *   - a thread parked on a consuming instruction that cannot take the next byte
 *     of the buffer is not appended (see pike_add_thread), and the first-byte
 *     prefilter (:256-309) becomes "the list holds only the .*? thread".
 * The thread records live in one block of global memory per context, the hot
 * rest (scalar state, marks, the top of the DFS stack, the working capture
 * window) in shared memory; the same routine serves the batch entry points
 * (context re-initialised per line; since the closure-table kernel of
 * sre_pike_table.cu took over the batch work this is its fallback: programs
 * whose tables do not fit, regexes with more than 15 groups, lines it gave up
 * on) and the streaming sre_vm_pike_exec (context persists between calls,
 * header and marks stored back into the block).
 */
#include <cstdlib>
#include "sre_kernels.cuh"

namespace {

enum {
    OP_CHAR = 1, OP_MATCH = 2, OP_JMP = 3, OP_SPLIT = 4, OP_ANY = 5, OP_SAVE = 6,
    OP_IN = 7, OP_NOTIN = 8, OP_ASSERT = 9
};
enum { AS_SMALL_Z = 0x01, AS_DOLLAR = 0x02, AS_BIG_B = 0x04, AS_SMALL_B = 0x08,
       AS_BIG_A = 0x10, AS_CARET = 0x20 };

constexpr int RC_DONE = -4;

struct pike_hdr_t {
    uint32_t  tag, prog_tag;
    int64_t   processed_bytes;
    int32_t   head[2], tail[2], count[2];   /* two thread lists              */
    int32_t   cur;                          /* which one is clist            */
    int32_t   free_head, pool_used;
    int32_t   has_matched, matched_id;
    int64_t   last_matched_pos;
    int64_t   pending[2];
    uint8_t   first_buf, eof, empty_capture, seen_newline, seen_word, error;
    uint8_t   seen_start_state;             /* faithful mode: sre_vm_pike.c:796-800 */
    uint8_t   pad;
    uint32_t  pad2[2];                      /* 26 words: 2-way bank conflicts at most as a shared array */
};
static_assert(sizeof(pike_hdr_t) == 104, "pike_hdr_t layout");

/*
 * Context memory.  Every array element is one or two 32-bit words; word e of a
 * context lives at base[e * stride].  Batch kernels interleave the contexts of
 * a warp (stride 32, lane l starts at word l), so that lanes touching the same
 * element -- the common case, they run the same program on similar lines --
 * share 128-byte lines in L1/L2; the streaming ctx of the classic API is a
 * single context with stride 1.  The scalar header lives in shared memory and
 * is only stored in the block between streaming calls.
 */
struct pike_ctx_t {
    pike_hdr_t   *h;            /* scalar state, in shared memory */
    uint32_t     *base;
    uint32_t      stride;
    uint32_t      o_matched, o_thr, o_stk, o_init, o_hdr;
    uint32_t      max_slots, rec;       /* rec = words of one thread record */
    /* working capture: only the slots [cap_base, cap_base + max_slots) of one
     * regex can be set at any time (see header), so that window is all that
     * is stored: 64-bit slot j at capp[(2j, 2j+1) * cstride] */
    uint32_t     *capp;
    uint32_t      cstride, cap_base;
    /* first SK entries of the DFS stack (shared memory in the kernels); deeper
     * entries spill to the context block */
    uint32_t     *stkp;
    uint32_t      sstride, sk;
    /* dedup marks: two bit sets (epoch parity) of tagw words each; word i of
     * set p is tagp[(p * tagw + i) * tstride] -- shared memory in the batch
     * kernels when it fits, else part of the context block */
    uint32_t     *tagp;
    uint32_t      tagw, tstride;

    __device__ __forceinline__ uint32_t &W(uint32_t e) const { return base[(size_t) e * stride]; }
    __device__ __forceinline__ int64_t get64(uint32_t e) const
    {
        return (int64_t) (((uint64_t) W(e + 1) << 32) | W(e));
    }
    __device__ __forceinline__ void set64(uint32_t e, int64_t v) const
    {
        W(e) = (uint32_t) v;
        W(e + 1) = (uint32_t) ((uint64_t) v >> 32);
    }
    __device__ __forceinline__ uint32_t &tagword(uint32_t tag, uint32_t i) const
    {
        return tagp[(size_t) ((tag & 1) * tagw + i) * tstride];
    }
    __device__ __forceinline__ bool tag_test(uint32_t pc, uint32_t tag) const
    {
        return (tagword(tag, pc >> 5) >> (pc & 31)) & 1;
    }
    __device__ __forceinline__ void tag_set(uint32_t pc, uint32_t tag) const
    {
        tagword(tag, pc >> 5) |= 1u << (pc & 31);
    }
    /* open epoch `tag`: forget the marks of epoch tag - 2 */
    __device__ __forceinline__ void tag_open(uint32_t tag) const
    {
        for (uint32_t i = 0; i < tagw; i++) {
            tagword(tag, i) = 0;
        }
    }
    /* thread record t: {pc, next, seen_word, cap[max_slots]} */
    __device__ __forceinline__ int32_t &t_pc(uint32_t t) const
    {
        return reinterpret_cast<int32_t &>(W(o_thr + t * rec));
    }
    __device__ __forceinline__ int32_t &t_next(uint32_t t) const
    {
        return reinterpret_cast<int32_t &>(W(o_thr + t * rec + 1));
    }
    __device__ __forceinline__ uint32_t &t_sw(uint32_t t) const { return W(o_thr + t * rec + 2); }
    /* absolute slot i of the working capture */
    __device__ __forceinline__ int64_t cap(uint32_t i) const
    {
        const uint32_t j = i - cap_base;
        if (j >= max_slots) {
            return -1;
        }
        return (int64_t) (((uint64_t) capp[(size_t) (2 * j + 1) * cstride] << 32) | capp[(size_t) 2 * j * cstride]);
    }
    /* i must lie in the current window */
    __device__ __forceinline__ void set_cap(uint32_t i, int64_t v) const
    {
        const uint32_t j = i - cap_base;
        if (j < max_slots) {
            capp[(size_t) 2 * j * cstride] = (uint32_t) v;
            capp[(size_t) (2 * j + 1) * cstride] = (uint32_t) ((uint64_t) v >> 32);
        }
    }
    __device__ __forceinline__ void cap_reset() const
    {
        for (uint32_t j = 0; j < 2 * max_slots; j++) {
            capp[(size_t) j * cstride] = 0xffffffffu;
        }
    }
    __device__ __forceinline__ int64_t matched(uint32_t i) const { return get64(o_matched + 2 * i); }
    __device__ __forceinline__ void set_matched(uint32_t i, int64_t v) const { set64(o_matched + 2 * i, v); }
    __device__ __forceinline__ int64_t t_cap(uint32_t t, uint32_t i) const
    {
        return get64(o_thr + t * rec + 3 + 2 * i);
    }
    __device__ __forceinline__ void set_t_cap(uint32_t t, uint32_t i, int64_t v) const
    {
        set64(o_thr + t * rec + 3 + 2 * i, v);
    }
    /* DFS stack entry: {kind, pc, val} */
    __device__ __forceinline__ uint32_t &stkw(uint32_t i, uint32_t k) const
    {
        if (i < sk) {
            return stkp[(size_t) (4 * i + k) * sstride];
        }
        return W(o_stk + 4 * i + k);
    }
    __device__ __forceinline__ int32_t stk_kind(uint32_t i) const { return (int32_t) stkw(i, 0); }
    __device__ __forceinline__ int32_t stk_pc(uint32_t i) const { return (int32_t) stkw(i, 1); }
    __device__ __forceinline__ int64_t stk_val(uint32_t i) const
    {
        return (int64_t) (((uint64_t) stkw(i, 3) << 32) | stkw(i, 2));
    }
    __device__ __forceinline__ void stk_set(uint32_t i, int32_t kind, int32_t pc, int64_t val) const
    {
        stkw(i, 0) = (uint32_t) kind;
        stkw(i, 1) = (uint32_t) pc;
        stkw(i, 2) = (uint32_t) val;
        stkw(i, 3) = (uint32_t) ((uint64_t) val >> 32);
    }
};

/* words of one context block; the same walk gives the section offsets */
struct pike_layout_t {
    uint32_t o_tags, o_matched, o_cap, o_thr, o_stk, o_init, o_hdr, tagw, rec;
    size_t   words;
};

__host__ __device__ inline pike_layout_t pike_layout(uint32_t len, uint32_t nslots, uint32_t max_slots,
                                                     uint32_t nthreads, uint32_t stack_cap)
{
    pike_layout_t L;
    size_t o = 0;
    L.tagw = (len + 1 + 31) / 32;
    L.rec = 3 + 2 * max_slots;
    L.o_tags = (uint32_t) o;     o += 2 * (size_t) L.tagw;
    L.o_matched = (uint32_t) o;  o += 2 * (size_t) nslots;
    L.o_cap = (uint32_t) o;      o += 2 * (size_t) max_slots;
    L.o_thr = (uint32_t) o;      o += (size_t) nthreads * L.rec;
    L.o_stk = (uint32_t) o;      o += 4 * (size_t) stack_cap;
    /* faithful mode: the initial thread list as the reference records it (:217-228): its
     * length, then the pcs of all its threads but the last */
    L.o_init = (uint32_t) o;     o += (size_t) len + 2;
    L.o_hdr = (uint32_t) o;      o += (sizeof(pike_hdr_t) + 3) / 4;    /* streaming ctx only */
    L.words = o;
    return L;
}

/* shared memory of a block of B contexts: [marks | stack | capture window],
 * each part optional (bit 0 / 1 / 2 of `parts`) */
enum { SM_MARKS = 1, SM_STACK = 2, SM_CAP = 4, SM_STACK_ENTRIES = 8 };

__host__ __device__ inline size_t pike_smem_words(const sre_dev_pike_t &pk, int parts, uint32_t B)
{
    size_t w = 0;
    if (parts & SM_MARKS) {
        w += (size_t) 2 * ((pk.len + 1 + 31) / 32) * B;
    }
    if (parts & SM_STACK) {
        w += (size_t) 4 * SM_STACK_ENTRIES * B;
    }
    if (parts & SM_CAP) {
        w += (size_t) 2 * pk.max_slots * B;
    }
    w = (w + 1) & ~(size_t) 1;
    w += sizeof(pike_hdr_t) / 4 * B;        /* always */
    return w;
}

/* attach a context view to its memory: base = first word of this lane of the
 * context block; smem = the block's shared memory, this context being lane
 * `lane` of B */
__device__ __forceinline__ void pike_attach(pike_ctx_t &c, const sre_dev_pike_t &pk, uint32_t *base, uint32_t stride,
                                   uint32_t *smem, int parts, uint32_t lane, uint32_t B)
{
    const pike_layout_t L = pike_layout(pk.len, pk.nslots, pk.max_slots, pk.max_threads, pk.stack_cap);
    uint32_t *const smem0 = smem;
    c.base = base;
    c.stride = stride;
    c.max_slots = pk.max_slots;
    c.rec = L.rec;
    c.o_matched = L.o_matched;
    c.o_thr = L.o_thr;
    c.o_stk = L.o_stk;
    c.o_init = L.o_init;
    c.o_hdr = L.o_hdr;
    c.tagw = L.tagw;
    c.cap_base = 0;
    if (parts & SM_MARKS) {
        c.tagp = smem + lane;
        c.tstride = B;
        smem += (size_t) 2 * L.tagw * B;
    } else {
        c.tagp = base + (size_t) L.o_tags * stride;
        c.tstride = stride;
    }
    if (parts & SM_STACK) {
        c.stkp = smem + lane;
        c.sstride = B;
        c.sk = SM_STACK_ENTRIES;
        smem += (size_t) 4 * SM_STACK_ENTRIES * B;
    } else {
        c.stkp = nullptr;
        c.sstride = 0;
        c.sk = 0;
    }
    if (parts & SM_CAP) {
        c.capp = smem + lane;
        c.cstride = B;
        smem += (size_t) 2 * pk.max_slots * B;
    } else {
        c.capp = base + (size_t) L.o_cap * stride;
        c.cstride = stride;
    }
    /* the scalar state is indexed dynamically (head[l] ...): as a local
     * variable it would live in local memory, so it is always shared */
    smem += (smem - smem0) & 1;
    c.h = reinterpret_cast<pike_hdr_t *>(smem) + lane;
}

/* batch layout: contexts 32g .. 32g+31 share one interleaved block */
__device__ __forceinline__ uint32_t *batch_base(const sre_dev_pike_t &pk, uint8_t *scratch, size_t tid, uint32_t stride)
{
    if (stride == 1) {
        return reinterpret_cast<uint32_t *>(scratch + tid * pk.ctx_stride);
    }
    return reinterpret_cast<uint32_t *>(scratch + (tid >> 5) * (32 * pk.ctx_stride)) + (tid & 31);
}

__device__ __forceinline__ void pike_hdr_store(const sre_dev_pike_t &pk, pike_ctx_t &c)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(c.h);
    const uint32_t o = c.o_hdr;
    for (uint32_t i = 0; i < (sizeof(pike_hdr_t) + 3) / 4; i++) {
        c.W(o + i) = src[i];
    }
}

__device__ __forceinline__ void pike_hdr_load(const sre_dev_pike_t &pk, pike_ctx_t &c)
{
    uint32_t *dst = reinterpret_cast<uint32_t *>(c.h);
    const uint32_t o = c.o_hdr;
    for (uint32_t i = 0; i < (sizeof(pike_hdr_t) + 3) / 4; i++) {
        dst[i] = c.W(o + i);
    }
}

__device__ __forceinline__ bool isword(uint32_t c)
{
    return (c - '0' < 10u) || ((c | 0x20) - 'a' < 26u) || c == '_';
}

__device__ __forceinline__ bool in_ranges(const sre_dev_pike_t &pk, const sre_dev_inst_t &in, uint32_t b)
{
    const uint8_t *r = pk.ranges + 2 * (size_t) in.v;
    for (uint32_t j = 0; j < in.nranges; j++) {
        if (b >= r[2 * j] && b <= r[2 * j + 1]) {
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ bool consumes(const sre_dev_pike_t &pk, const sre_dev_inst_t &in, uint32_t b)
{
    switch (in.opcode) {
    case OP_CHAR:  return in.ch == b;
    case OP_ANY:   return true;
    case OP_IN:    return in_ranges(pk, in, b);
    case OP_NOTIN: return !in_ranges(pk, in, b);
    default:       return false;
    }
}

/* fresh context for a new stream / line */
__device__ __forceinline__ void pike_reset(pike_ctx_t &c, bool first_time)
{
    pike_hdr_t *h = c.h;
    if (first_time) {
        h->tag = 0;
        h->prog_tag = 0;
    }
    h->processed_bytes = 0;
    h->head[0] = h->head[1] = -1;
    h->tail[0] = h->tail[1] = -1;
    h->count[0] = h->count[1] = 0;
    h->cur = 0;
    h->free_head = -1;
    h->pool_used = 0;
    h->has_matched = 0;
    h->matched_id = 0;
    h->last_matched_pos = -1;
    h->first_buf = 1;
    h->eof = 0;
    h->empty_capture = 0;
    h->seen_newline = 0;
    h->seen_word = 0;
    h->error = 0;
    h->seen_start_state = 0;
}

__device__ __forceinline__ void list_clear(pike_ctx_t &c, int l)
{
    pike_hdr_t *h = c.h;
    if (h->head[l] >= 0) {
        c.t_next(h->tail[l]) = h->free_head;
        h->free_head = h->head[l];
    }
    h->head[l] = h->tail[l] = -1;
    h->count[l] = 0;
}

/* working capture <- thread t's slots (everything else is -1 already) */
__device__ __forceinline__ void cap_load(const sre_dev_pike_t &pk, pike_ctx_t &c, int32_t t, uint32_t *base_out,
                                uint32_t *cnt_out)
{
    const uint32_t r = pk.pc_regex[c.t_pc(t)], base = pk.slot_ofs[r];
    const uint32_t cnt = pk.slot_ofs[r + 1] - base;
    c.cap_base = base;
    for (uint32_t i = 0; i < cnt; i++) {
        c.set_cap(base + i, c.t_cap(t, i));
    }
    *base_out = base;
    *cnt_out = cnt;
}

__device__ __forceinline__ void cap_clear(pike_ctx_t &c, uint32_t base, uint32_t cnt)
{
    /* add_thread may have moved the (then all -1) window to another regex */
    c.cap_base = base;
    for (uint32_t i = 0; i < cnt; i++) {
        c.set_cap(base + i, -1);
    }
}

/* slot g of thread t's full capture vector */
__device__ __forceinline__ int64_t thread_slot(const sre_dev_pike_t &pk, const pike_ctx_t &c, int32_t t, uint32_t g)
{
    const uint32_t r = pk.pc_regex[c.t_pc(t)], base = pk.slot_ofs[r];
    if (g < base || g >= pk.slot_ofs[r + 1]) {
        return -1;
    }
    return c.t_cap(t, g - base);
}

/* a temporary list used by assertion_hold */
struct tmp_list_t { int32_t head, tail, count; };

/* take a thread record, park it on pc and append it to list l (or *tmp);
 * the caller fills its capture.  -1: pool exhausted */
__device__ __forceinline__ int32_t thread_append(const sre_dev_pike_t &pk, pike_ctx_t &c, int l, tmp_list_t *tmp,
    int32_t pc, uint32_t seen_word)
{
    pike_hdr_t *h = c.h;
    int32_t t;
    if (h->free_head >= 0) {
        t = h->free_head;
        h->free_head = c.t_next(t);
    } else if (h->pool_used < (int32_t) pk.max_threads) {
        t = h->pool_used++;
    } else {
        return -1;
    }
    c.t_pc(t) = pc;
    c.t_sw(t) = seen_word;
    c.t_next(t) = -1;
    if (l >= 0) {
        if (h->head[l] < 0) {
            h->head[l] = t;
        } else {
            c.t_next(h->tail[l]) = t;
        }
        h->tail[l] = t;
        h->count[l]++;
    } else {
        if (tmp->head < 0) {
            tmp->head = t;
        } else {
            c.t_next(tmp->tail) = t;
        }
        tmp->tail = t;
        tmp->count++;
    }
    return t;
}

/*
 * add_thread (sre_vm_pike.c:756-942).  Appends to list `l` (0/1) or, when
 * l < 0, to *tmp.  c.cap holds the capture of the calling thread and is
 * restored to that value on return.  Returns SRE_K_OK, RC_DONE or SRE_K_ERROR.
 *
 * nb = the byte the new threads will be stepped on (0..255), NB_END when they
 * will only see the end of input, NB_UNKNOWN when it is not in this buffer.
 * A thread parked on a consuming instruction that cannot take nb would be
 * dropped by the next step without any effect, so it is not appended at all;
 * for the same reason the whole closure of the regexes is skipped at the
 * start state when nb is outside the leading-byte set (nleading != 0 means
 * every path from pc 0 must first consume a leading byte).  Lists at the end
 * of a non-final buffer are complete (NB_UNKNOWN), so AGAIN, the temporary
 * captures and pending matches are those of the reference.
 */
enum { NB_END = -2, NB_UNKNOWN = -1 };

__device__ __forceinline__ int pike_add_thread(const sre_dev_pike_t &pk, pike_ctx_t &c, int l, tmp_list_t *tmp,
    int32_t pc0, int64_t pos, const uint8_t *buffer, bool want_done, int nb, bool faithful = false)
{
    pike_hdr_t *h = c.h;
    const uint32_t tag = h->tag;
    int sp = 0;
    c.stk_set(sp, -1, pc0, 0);
    sp++;

    while (sp > 0) {
        sp--;
        if (c.stk_kind(sp) >= 0) {
            c.set_cap(c.stk_kind(sp), c.stk_val(sp));
            continue;
        }
        int32_t pc = c.stk_pc(sp);

        for (;;) {
            const sre_dev_inst_t in = pk.insts[pc];
            uint32_t seen_word = 0;
            bool add = false;

            if (c.tag_test(pc, tag)) {
                /* the revisited-SPLIT rule, :770-786 */
                if (in.opcode == OP_SPLIT && !c.tag_test(in.y, tag)) {
                    if (faithful && pc == 0) {
                        h->seen_start_state = 1;        /* :775-778 */
                    }
                    pc = in.y;
                    continue;
                }
                break;
            }
            c.tag_set(pc, tag);
            if (faithful && pc == 0 && in.opcode == OP_SPLIT) {
                h->seen_start_state = 1;                /* :797-800 */
            }

            switch (in.opcode) {
            case OP_JMP:
                pc = in.x;
                continue;

            case OP_SPLIT:
                if (pc == 0 && pk.nleading && nb != NB_UNKNOWN) {
                    if (nb == NB_END || !((pk.leadset[nb >> 5] >> (nb & 31)) & 1)) {
                        pc = in.y;      /* only the ".*?" thread can go on */
                        continue;
                    }
                    if (pk.start_ofs) {
                        /* the precomputed start closure for this byte; the
                         * working capture is all -1 here (pc 0 is only reached
                         * from the start and from the ".*?" thread) */
                        const uint32_t e1 = pk.start_ofs[nb + 1];
                        for (uint32_t e = pk.start_ofs[nb]; e < e1; e++) {
                            const sre_dev_start_t se = pk.start_ent[e];
                            if (c.tag_test(se.pc, tag)) {
                                continue;
                            }
                            c.tag_set(se.pc, tag);
                            const int32_t t = thread_append(pk, c, l, tmp, se.pc, 0);
                            if (t < 0) {
                                return SRE_K_ERROR;
                            }
                            const uint32_t r = pk.pc_regex[se.pc];
                            const uint32_t cnt = pk.slot_ofs[r + 1] - pk.slot_ofs[r];
                            for (uint32_t i = 0; i < cnt; i++) {
                                c.set_t_cap(t, i, -1);
                            }
                            for (uint32_t k = 0; k < se.nsl; k++) {
                                c.set_t_cap(t, se.sl[k], h->processed_bytes + pos);
                            }
                        }
                        pc = in.y;
                        continue;
                    }
                }
                if (sp >= (int) pk.stack_cap) {
                    return SRE_K_ERROR;
                }
                c.stk_set(sp, -1, in.y, 0);
                sp++;
                pc = in.x;
                continue;

            case OP_SAVE:
                if (sp >= (int) pk.stack_cap) {
                    return SRE_K_ERROR;
                }
                c.cap_base = pk.slot_ofs[pk.pc_regex[pc]];
                c.stk_set(sp, in.v, 0, c.cap(in.v));
                sp++;
                c.set_cap(in.v, h->processed_bytes + pos);
                pc++;
                continue;

            case OP_ASSERT:
                switch (in.v) {
                case AS_BIG_A:
                    if (pos || h->processed_bytes) {
                        break;
                    }
                    pc++;
                    continue;
                case AS_CARET:
                    if (pos == 0) {
                        if (h->processed_bytes && !h->seen_newline) {
                            break;
                        }
                    } else if (buffer[pos - 1] != '\n') {
                        break;
                    }
                    pc++;
                    continue;
                case AS_SMALL_B:
                case AS_BIG_B:
                    seen_word = pos == 0 ? 0 : isword(buffer[pos - 1]);
                    add = true;
                    break;
                default:            /* look-ahead: postponed */
                    add = true;
                    break;
                }
                break;

            case OP_MATCH:
                h->last_matched_pos = c.cap(1);
                if (want_done) {
                    for (uint32_t i = 0; i < pk.nslots; i++) {
                        c.set_matched(i, c.cap(i));
                    }
                    h->matched_id = in.v;
                    /* unwind the undo log so c.cap is the caller's again */
                    while (sp > 0) {
                        sp--;
                        if (c.stk_kind(sp) >= 0) {
                            c.set_cap(c.stk_kind(sp), c.stk_val(sp));
                        }
                    }
                    return RC_DONE;
                }
                add = true;
                break;

            default:                    /* CHAR / ANY / IN / NOTIN */
                add = !(nb == NB_END || (nb >= 0 && !consumes(pk, in, (uint32_t) nb)));
                break;
            }

            if (add) {
                const int32_t t = thread_append(pk, c, l, tmp, pc, seen_word);
                if (t < 0) {
                    return SRE_K_ERROR;
                }
                /* only the owning regex's slots can be set (see header) */
                const uint32_t r = pk.pc_regex[pc], base = pk.slot_ofs[r];
                const uint32_t cnt = pk.slot_ofs[r + 1] - base;
                for (uint32_t i = 0; i < cnt; i++) {
                    c.set_t_cap(t, i, c.cap(base + i));
                }
            }
            break;
        }
    }
    return SRE_K_OK;
}

__device__ __forceinline__ int64_t find_first_byte(const sre_dev_pike_t &pk, const uint8_t *buf, int64_t pos,
    int64_t last)
{
    if (pk.leading_byte != -1) {
        const uint32_t lb = (uint32_t) pk.leading_byte;
        while (pos < last && buf[pos] != lb) {
            pos++;
        }
        return pos;
    }
    for (; pos != last; pos++) {
        const uint32_t b = buf[pos];
        for (uint32_t i = 0; i < pk.nleading; i++) {
            if (consumes(pk, pk.insts[pk.leading[i]], b)) {
                return pos;
            }
        }
    }
    return pos;
}

/* prepare_matched_captures, :945-989.  complete: whole slice + -1 fill */
__device__ __forceinline__ int prepare_matched(const sre_dev_pike_t &pk, pike_ctx_t &c, int64_t *ovector,
    uint32_t ovec_slots, bool complete)
{
    const int32_t id = c.h->matched_id;
    if (id < 0 || (uint32_t) id >= pk.nregexes) {
        return SRE_K_ERROR;
    }
    const uint32_t ofs = pk.slot_ofs[id];
    const uint32_t n = complete ? pk.slot_ofs[id + 1] - ofs : 2;
    for (uint32_t i = 0; i < n && i < ovec_slots; i++) {
        ovector[i] = c.matched(ofs + i);
    }
    if (complete) {
        for (uint32_t i = n; i < ovec_slots; i++) {
            ovector[i] = -1;
        }
    }
    return SRE_K_OK;
}

/*
 * One sre_vm_pike_exec call (:148-689).  ovector: caller's vector
 * (ovec_slots entries).  *pending_set: 1 when h->pending holds a pending match.
 */
/*
 * faithful: the reference's first-byte prefilter to the letter (:256-309) over complete thread
 * lists (no next-byte pruning), including its misfire after a match (the "is this the initial
 * list" test compares the count and every pc but the last, :262-274).  The batch
 * tiers run without it (the prefilter below is result neutral) and hand the lines on which the
 * misfire is possible to a faithful pass (k_pike_quirk_mark).
 */
__device__ __forceinline__ int pike_exec(const sre_dev_pike_t &pk, pike_ctx_t &c, const uint8_t *input, int64_t size,
    bool eof, int64_t *ovector, uint32_t ovec_slots, int *pending_set, int64_t start_pos = 0,
    bool faithful = false)
{
    pike_hdr_t *h = c.h;
    int64_t sp, last = size;
    int rc;

    *pending_set = 0;
    if (h->eof) {
        return SRE_K_ERROR;
    }
    h->last_matched_pos = -1;

    if (h->empty_capture) {                         /* :179-193 */
        h->empty_capture = 0;
        if (size == 0) {
            if (eof) {
                h->eof = 1;
                return SRE_K_DECLINED;
            }
            return SRE_K_AGAIN;
        }
        sp = 1;
    } else {
        /* batch callers may pass an offset before which no match can start
         * (see k_dfa_lines_hint); the classic API always starts at 0 */
        sp = start_pos;
    }

    int cl = h->cur, nl = cl ^ 1;

    if (h->first_buf) {                             /* :202-229 */
        h->first_buf = 0;
        c.cap_base = 0;
        c.cap_reset();
        h->tag = h->prog_tag + 1;
        c.tag_open(h->tag);
        rc = pike_add_thread(pk, c, cl, nullptr, 0, sp, input, false,
                             faithful ? NB_UNKNOWN : sp < last ? (int) input[sp] : (eof ? NB_END : NB_UNKNOWN),
                             faithful);
        if (rc != SRE_K_OK) {
            h->prog_tag = h->tag;
            return SRE_K_ERROR;
        }
        if (faithful) {                             /* :217-228 */
            c.W(c.o_init) = (uint32_t) h->count[cl];
            uint32_t i = 0;
            for (int32_t t = h->head[cl]; t >= 0 && c.t_next(t) >= 0; t = c.t_next(t)) {
                c.W(c.o_init + 1 + i) = (uint32_t) c.t_pc(t);
                i++;
            }
        }
    } else {
        h->tag = h->prog_tag;
    }

    for (; sp < last || (eof && sp == last); sp++) {
        if (h->head[cl] < 0) {
            break;
        }

        /*
         * First-byte prefilter (:256-309, find_first_byte :992-1061; result-
         * neutral in the reference too).  Here the list holds nothing but the
         * ".*?" thread exactly when no regex thread is in flight, and then the
         * scan can move to just before the next byte a leading instruction
         * takes: stepping the ".*?" thread there rebuilds the start closure.
         */
        if (faithful) {
            if (pk.nleading && h->seen_start_state) {
                h->seen_start_state = 0;
                bool initial = sp != last && (uint32_t) h->count[cl] == c.W(c.o_init);
                if (initial) {
                    uint32_t i = 0;
                    for (int32_t t = h->head[cl]; t >= 0 && c.t_next(t) >= 0; t = c.t_next(t), i++) {
                        if ((uint32_t) c.t_pc(t) != c.W(c.o_init + 1 + i)) {
                            initial = false;
                            break;
                        }
                    }
                }
                if (initial) {
                    const int64_t p = find_first_byte(pk, input, sp, last);
                    if (p > sp) {
                        sp = p;
                        list_clear(c, cl);
                        c.cap_base = 0;
                        c.cap_reset();
                        h->tag++;
                        c.tag_open(h->tag);
                        rc = pike_add_thread(pk, c, cl, nullptr, 0, sp, input, false, NB_UNKNOWN, true);
                        if (rc != SRE_K_OK) {
                            h->prog_tag = h->tag;
                            return SRE_K_ERROR;
                        }
                        if (sp == last) {
                            break;
                        }
                    }
                }
            }
        } else
        if (pk.nleading && h->count[cl] == 1 && sp + 1 < last && c.t_pc(h->head[cl]) == 1) {
            const int64_t p = find_first_byte(pk, input, sp + 1, last);
            if (p - 1 > sp) {
                sp = p - 1;
            }
        }

        h->tag++;
        c.tag_open(h->tag);
        const bool at_end = (sp == last);
        const uint32_t byte = at_end ? 0 : input[sp];
        const bool cur_word = !at_end && isword(byte);
        const int nb_cur = faithful ? NB_UNKNOWN : at_end ? NB_END : (int) byte;
        const int nb_next = faithful ? NB_UNKNOWN : sp + 1 < last ? (int) input[sp + 1] : (eof ? NB_END : NB_UNKNOWN);

        while (h->head[cl] >= 0) {                  /* :314-567 */
            const int32_t t = h->head[cl];
            h->head[cl] = c.t_next(t);
            if (h->head[cl] < 0) {
                h->tail[cl] = -1;
            }
            h->count[cl]--;

            const int32_t pc = c.t_pc(t);
            const sre_dev_inst_t in = pk.insts[pc];
            bool got_match = false;
            uint32_t cbase = 0, ccnt = 0;

            if (in.opcode == OP_ASSERT) {           /* :449-528 */
                const bool seen_word = c.t_sw(t) != 0 || (sp == 0 && h->seen_word);
                bool hold = false;
                switch (in.v) {
                case AS_SMALL_Z: hold = at_end; break;
                case AS_DOLLAR:  hold = at_end || byte == '\n'; break;
                case AS_BIG_B:   hold = (seen_word == cur_word); break;
                case AS_SMALL_B: hold = (seen_word != cur_word); break;
                default: break;
                }
                if (hold) {
                    cap_load(pk, c, t, &cbase, &ccnt);
                    tmp_list_t tl = { -1, -1, 0 };
                    h->tag--;
                    rc = pike_add_thread(pk, c, -1, &tl, pc + 1, sp, input, false, nb_cur, faithful);
                    if (rc != SRE_K_OK) {
                        h->prog_tag = h->tag + 1;
                        return SRE_K_ERROR;
                    }
                    if (tl.head >= 0) {             /* prepend, :519-523 */
                        c.t_next(tl.tail) = h->head[cl];
                        if (h->head[cl] < 0) {
                            h->tail[cl] = tl.tail;
                        }
                        h->head[cl] = tl.head;
                        h->count[cl] += tl.count;
                    }
                    h->tag++;
                    cap_clear(c, cbase, ccnt);
                }
            } else if (in.opcode == OP_MATCH) {     /* :530-553 */
                cap_load(pk, c, t, &cbase, &ccnt);
                h->last_matched_pos = c.cap(1);
                for (uint32_t i = 0; i < pk.nslots; i++) {
                    c.set_matched(i, c.cap(i));
                }
                h->matched_id = in.v;
                got_match = true;
                cap_clear(c, cbase, ccnt);
            } else if (!at_end && consumes(pk, in, byte)) {
                cap_load(pk, c, t, &cbase, &ccnt);
                rc = pike_add_thread(pk, c, nl, nullptr, pc + 1, sp + 1, input, true, nb_next, faithful);
                cap_clear(c, cbase, ccnt);
                if (rc == RC_DONE) {
                    got_match = true;
                } else if (rc != SRE_K_OK) {
                    h->prog_tag = h->tag;
                    return SRE_K_ERROR;
                }
            }

            /* free the thread */
            c.t_next(t) = h->free_head;
            h->free_head = t;

            if (got_match) {
                h->has_matched = 1;
                list_clear(c, cl);
                break;
            }
        }

        /* step_done: swap lists */
        cl ^= 1;
        nl ^= 1;
        list_clear(c, nl);
        if (at_end) {
            break;
        }
    }

    if (h->last_matched_pos >= 0) {                 /* :586-601 */
        const int64_t p = h->last_matched_pos - h->processed_bytes;
        if (p > 0) {
            h->seen_newline = (input[p - 1] == '\n');
            h->seen_word = isword(input[p - 1]);
        }
        h->last_matched_pos = -1;
    }

    h->prog_tag = h->tag;
    h->cur = cl;

    if (h->has_matched) {
        if (eof || h->head[cl] < 0) {               /* :607-636 */
            if (prepare_matched(pk, c, ovector, ovec_slots, true) != SRE_K_OK) {
                return SRE_K_ERROR;
            }
            if (h->head[cl] >= 0) {
                list_clear(c, cl);
                h->eof = 1;
            }
            h->processed_bytes = ovector[1];
            h->empty_capture = (ovector[0] == ovector[1]);
            h->has_matched = 0;
            h->first_buf = 1;
            return h->matched_id;
        }
        /* :640-658 */
        *pending_set = 1;
        if (prepare_matched(pk, c, h->pending, 2, false) != SRE_K_OK) {
            return SRE_K_ERROR;
        }
    } else if (eof) {
        h->eof = 1;
        return SRE_K_DECLINED;
    }

    h->processed_bytes += sp;

    /* prepare_temp_captures, :692-735 (note: end slot read without the
     * per-regex offset, :721) */
    ovector[0] = -1;
    if (ovec_slots > 1) {
        ovector[1] = -1;
    }
    for (int32_t t = h->head[cl]; t >= 0; t = c.t_next(t)) {
        for (uint32_t i = 0; i < pk.nregexes; i++) {
            const int64_t b0 = thread_slot(pk, c, t, pk.slot_ofs[i]);
            if (b0 != -1 && (ovector[0] == -1 || b0 < ovector[0])) {
                ovector[0] = b0;
            }
            const int64_t b1 = thread_slot(pk, c, t, 1);
            if (ovec_slots > 1 && b1 != -1 && (ovector[1] == -1 || b1 > ovector[1])) {
                ovector[1] = b1;
            }
        }
    }
    return SRE_K_AGAIN;
}

/* ---- kernels --------------------------------------------------------------- */

/* the lines the gate let through, packed, so that every lane of the Pike
 * kernels has work (warp-aggregated append; the order is irrelevant) */
__global__ void __launch_bounds__(256)
k_pike_compact(const int32_t *__restrict__ select, size_t nlines, int32_t *__restrict__ rc,
               uint32_t *__restrict__ list, uint32_t *__restrict__ count)
{
    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    const size_t rounds = (nlines + nthreads - 1) / nthreads;
    const uint32_t lane = threadIdx.x & 31;
    for (size_t r = 0; r < rounds; r++) {
        const size_t line = r * nthreads + (size_t) blockIdx.x * blockDim.x + threadIdx.x;
        const int32_t sel = line < nlines ? select[line] : SRE_K_DECLINED;
        const bool take = line < nlines && sel == SRE_K_OK;
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        uint32_t base = 0;
        if (lane == 0 && m) {
            base = atomicAdd(count, (uint32_t) __popc(m));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) {
            list[base + __popc(m & ((1u << lane) - 1))] = (uint32_t) line;
        } else if (line < nlines) {
            rc[line] = sel;
        }
    }
}

/* see sre_launch_pike_quirk_mark (sre_kernels.cuh) */
__global__ void __launch_bounds__(256)
k_pike_quirk_mark(sre_dev_pike_t pk, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
                  size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines, int32_t *__restrict__ rc,
                  const int64_t *__restrict__ ovec, uint32_t ovec_slots, uint32_t *__restrict__ count)
{
    const size_t nwork = lines.list ? (size_t) *lines.count : nlines;
    auto bit = [](const uint32_t *set, uint32_t b) { return (set[b >> 5] >> (b & 31)) & 1u; };
    uint32_t mine = 0;
    for (size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < nwork; k += (size_t) gridDim.x * blockDim.x) {
        const size_t line = lines.list ? (size_t) lines.list[k] : k;
        if (rc[line] < 0) {
            continue;
        }
        bool cand = true;
        if (ovec != nullptr && ovec_slots >= 1) {
            const size_t start = offsets ? (size_t) offsets[line] : line * pitch;
            const int64_t size = (int64_t) (offsets ? (size_t) offsets[line + 1] - start : linelen);
            const int64_t s = ovec[line * ovec_slots];
            cand = s >= 1 && s + 1 < size && bit(pk.quirk_single, buf[start + s])
                   && !bit(pk.leadset, buf[start + s - 1]) && !bit(pk.leadset, buf[start + s + 1]);
        }
        if (cand) {
            rc[line] = SRE_K_QUIRK;
            mine++;
        }
    }
    if (mine) {
        atomicAdd(count, mine);
    }
}

__global__ void __launch_bounds__(128)
k_pike_lines(sre_dev_pike_t pk, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
             size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines,
             const int32_t *__restrict__ start_hint, int32_t *__restrict__ rc, int64_t *__restrict__ ovec,
             uint32_t ovec_slots, uint8_t *scratch, size_t nctx, int retry_only, uint32_t stride,
             int parts, const uint32_t *retry_count)
{
    if (retry_only && retry_count && *retry_count == 0) {
        return;                 /* the faster tiers gave up on no line */
    }
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nctx) {
        return;
    }
    extern __shared__ uint32_t smem_ctx[];
    pike_ctx_t c;
    pike_attach(c, pk, batch_base(pk, scratch, tid, stride), stride, smem_ctx, parts, threadIdx.x, blockDim.x);
    bool first = true;

    const size_t nwork = lines.list ? (size_t) *lines.count : nlines;
    for (size_t k = tid; k < nwork; k += nctx) {
        const size_t line = lines.list ? (size_t) lines.list[k] : k;
        int64_t *ov = ovec + line * ovec_slots;
        /* second pass after k_pike_small: only the lines it gave up on */
        if (retry_only && rc[line] != (retry_only == 2 ? SRE_K_QUIRK : SRE_K_RETRY)) {
            continue;
        }
        const size_t start = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t end = offsets ? (size_t) offsets[line + 1] : start + linelen;
        pike_reset(c, first);
        first = false;
        int pending;
        /* the quirk pass (retry_only 2) replays the reference to the letter from offset 0 */
        const int r = pike_exec(pk, c, buf + start, (int64_t) (end - start), true, ov, ovec_slots,
                                &pending, retry_only != 2 && start_hint ? start_hint[line] : 0, retry_only == 2);
        rc[line] = r;
        if (r < 0) {
            for (uint32_t i = 0; i < ovec_slots; i++) {
                ov[i] = -1;
            }
        }
    }
}

/*
 * All non-overlapping matches of every line: the reference's post-match
 * continuation (sre_vm_pike.c:624-635: after a match the ctx restarts at
 * ovector[1], and skips one byte after an empty match, :179-193), i.e. what
 * ngx_replace_filter does for global substitution, as one batch call.
 * spans[line][k] = ($0 start, $0 end) absolute in the line, ids[line][k] =
 * regex id, count[line] = matches found (capped at max_matches).
 */
__global__ void __launch_bounds__(128)
k_pike_lines_all(sre_dev_pike_t pk, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
                 size_t nlines, size_t pitch, size_t linelen, uint32_t max_matches,
                 int32_t *__restrict__ count, int64_t *__restrict__ spans, int32_t *__restrict__ ids,
                 uint8_t *scratch, size_t nctx, uint32_t stride, int parts)
{
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nctx) {
        return;
    }
    extern __shared__ uint32_t smem_ctx[];
    pike_ctx_t c;
    pike_attach(c, pk, batch_base(pk, scratch, tid, stride), stride, smem_ctx, parts, threadIdx.x, blockDim.x);
    bool first = true;
    for (size_t line = tid; line < nlines; line += nctx) {
        const size_t start = offsets ? (size_t) offsets[line] : line * pitch;
        const int64_t len = (int64_t) (offsets ? (size_t) offsets[line + 1] - start : linelen);
        pike_reset(c, first);
        first = false;
        uint32_t m = 0;
        int64_t at = 0, ov[2];
        while (m < max_matches) {
            int pending;
            const int r = pike_exec(pk, c, buf + start + at, len - at, true, ov, 2, &pending, 0, true);
            if (r < 0) {
                break;
            }
            spans[(line * max_matches + m) * 2] = ov[0];
            spans[(line * max_matches + m) * 2 + 1] = ov[1];
            ids[line * max_matches + m] = r;
            m++;
            at = ov[1];
            if (c.h->eof) {
                break;
            }
        }
        count[line] = (int32_t) m;
    }
}

__global__ void k_pike_ctx_init(sre_dev_pike_t pk, uint8_t *ctx)
{
    extern __shared__ uint32_t smem_ctx[];
    pike_ctx_t c;
    pike_attach(c, pk, reinterpret_cast<uint32_t *>(ctx), 1, smem_ctx, 0, 0, 1);
    c.tag_open(0);
    c.tag_open(1);
    pike_reset(c, true);
    pike_hdr_store(pk, c);
}

/* out[0] = rc, out[1] = pending flag, out[2..3] = pending, out[4..] = ovector */
__global__ void k_pike_stream(sre_dev_pike_t pk, uint8_t *ctx, const uint8_t *buf, size_t len, size_t skip,
                              int eof, int64_t *out, uint32_t ovec_slots, int parts)
{
    extern __shared__ uint32_t smem_ctx[];
    pike_ctx_t c;
    pike_attach(c, pk, reinterpret_cast<uint32_t *>(ctx), 1, smem_ctx, parts, 0, 1);
    pike_hdr_load(pk, c);
    c.cap_reset();      /* between calls the working capture is all -1 */
    if (skip > 0 && skip <= len && c.h->first_buf && c.h->processed_bytes == 0) {
        /* the state a context is in after `skip` bytes that left no thread alive: the offsets
         * go on from there, and the look-behind assertions of the threads added at the new
         * offset 0 see the byte before it (sre_vm_pike.c:839-887 read these fields) */
        c.h->processed_bytes = (int64_t) skip;
        c.h->seen_newline = buf[skip - 1] == '\n';
        c.h->seen_word = isword(buf[skip - 1]);
        buf += skip;
        len -= skip;
    }
    int pending = 0;
    const int r = pike_exec(pk, c, buf, (int64_t) len, eof != 0, out + 4, ovec_slots, &pending, 0, true);
    pike_hdr_store(pk, c);
    out[0] = r;
    out[1] = pending;
    out[2] = c.h->pending[0];
    out[3] = c.h->pending[1];
}

/* a batch of persistent contexts, one thread each: stream i is fed buf[off[i], off[i+1]);
 * out row i = { rc, pending flag, pending[0], pending[1], ovector... } as k_pike_stream's */
__global__ void __launch_bounds__(64)
k_pike_streams(sre_dev_pike_t pk, uint8_t *ctxs, size_t nstreams, const uint8_t *__restrict__ buf,
               const int64_t *__restrict__ off, const uint8_t *__restrict__ eofs, int eof_all, int64_t *out,
               uint32_t ovec_slots, int parts)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstreams) {
        return;
    }
    extern __shared__ uint32_t smem_ctx[];
    pike_ctx_t c;
    pike_attach(c, pk, reinterpret_cast<uint32_t *>(ctxs + i * pk.ctx_stride), 1, smem_ctx, parts, threadIdx.x,
                blockDim.x);
    pike_hdr_load(pk, c);
    c.cap_reset();
    int64_t *o = out + i * (4 + (size_t) ovec_slots);
    int pending = 0;
    const bool eof = eof_all || (eofs != nullptr && eofs[i] != 0);
    const int r = pike_exec(pk, c, buf + off[i], off[i + 1] - off[i], eof, o + 4, ovec_slots, &pending, 0, true);
    pike_hdr_store(pk, c);
    o[0] = r;
    o[1] = pending;
    o[2] = c.h->pending[0];
    o[3] = c.h->pending[1];
}

__global__ void __launch_bounds__(64)
k_pike_streams_init(sre_dev_pike_t pk, uint8_t *ctxs, size_t nstreams)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstreams) {
        return;
    }
    extern __shared__ uint32_t smem_ctx[];
    pike_ctx_t c;
    pike_attach(c, pk, reinterpret_cast<uint32_t *>(ctxs + i * pk.ctx_stride), 1, smem_ctx, 0, threadIdx.x, blockDim.x);
    c.tag_open(0);
    c.tag_open(1);
    pike_reset(c, true);
    pike_hdr_store(pk, c);
}

}  // namespace

/* per-context (1, default) or warp-interleaved (32) scratch; SRE_PIKE_INTERLEAVE=1
 * selects the latter (measured slower on divergent multi-pattern sets) */
static uint32_t batch_stride()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SRE_PIKE_INTERLEAVE");
        v = (e && atoi(e) > 0) ? 32 : 1;
    }
    return (uint32_t) v;
}

/* which parts of a context go to shared memory for blocks of B contexts;
 * *bytes = dynamic shared memory to ask for */
static int smem_parts(const sre_dev_pike_t &pk, uint32_t B, bool marks_allowed, size_t *bytes)
{
    static bool opted = false;
    if (!opted) {
        opted = true;
        cudaFuncSetAttribute(k_pike_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(k_pike_lines_all, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    }
    const size_t budget = B == 1 ? 40 * 1024 : 96 * 1024;
    int parts = SM_STACK;
    if (pike_smem_words(pk, parts | SM_CAP, B) * 4 <= budget / 2) {
        parts |= SM_CAP;
    }
    if (marks_allowed && pike_smem_words(pk, parts | SM_MARKS, B) * 4 <= budget) {
        parts |= SM_MARKS;
    }
    *bytes = pike_smem_words(pk, parts, B) * 4;
    return parts;
}

size_t sre_pike_ctx_bytes(uint32_t len, uint32_t nslots, uint32_t max_slots, uint32_t nthreads,
    uint32_t stack_cap)
{
    return pike_layout(len, nslots, max_slots, nthreads, stack_cap).words * 4;
}

cudaError_t sre_launch_pike_compact(const int32_t *select, size_t nlines, int32_t *rc, uint32_t *list,
    uint32_t *count, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    size_t grid = (nlines + 255) / 256;
    if (grid > 148 * 8) {
        grid = 148 * 8;
    }
    k_pike_compact<<<(unsigned) grid, 256, 0, stream>>>(select, nlines, rc, list, count);
    return cudaGetLastError();
}

cudaError_t sre_launch_pike_quirk_mark(const sre_dev_pike_t &pk, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines, int32_t *rc, const int64_t *ovec,
    uint32_t ovec_slots, uint32_t *count, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    cudaError_t err = cudaMemsetAsync(count, 0, 4, stream);
    if (err != cudaSuccess) {
        return err;
    }
    size_t grid = (nlines + 255) / 256;
    if (grid > 148 * 8) {
        grid = 148 * 8;
    }
    k_pike_quirk_mark<<<(unsigned) grid, 256, 0, stream>>>(pk, buf, offsets, nlines, pitch, linelen, lines, rc, ovec,
                                                          ovec_slots, count);
    return cudaGetLastError();
}

cudaError_t sre_launch_pike_lines(const sre_dev_pike_t &pk, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines,
    const int32_t *start, int32_t *rc, int64_t *ovec, uint32_t ovec_slots, uint8_t *scratch, size_t nctx,
    int retry_only, cudaStream_t stream, int *launches, const uint32_t *retry_count)
{
    if (nlines == 0 || nctx == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    const unsigned grid = (unsigned) ((nctx + 127) / 128);
    size_t smem;
    const int parts = smem_parts(pk, 128, true, &smem);
    k_pike_lines<<<grid, 128, smem, stream>>>(pk, buf, offsets, nlines, pitch, linelen, lines, start, rc,
                                             ovec, ovec_slots, scratch, nctx, retry_only, batch_stride(),
                                             parts, retry_count);
    return cudaGetLastError();
}

cudaError_t sre_launch_pike_lines_all(const sre_dev_pike_t &pk, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, uint32_t max_matches, int32_t *count, int64_t *spans,
    int32_t *ids, uint8_t *scratch, size_t nctx, cudaStream_t stream, int *launches)
{
    if (nlines == 0 || nctx == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    const unsigned grid = (unsigned) ((nctx + 127) / 128);
    size_t smem;
    const int parts = smem_parts(pk, 128, true, &smem);
    k_pike_lines_all<<<grid, 128, smem, stream>>>(pk, buf, offsets, nlines, pitch, linelen, max_matches,
                                                 count, spans, ids, scratch, nctx, batch_stride(), parts);
    return cudaGetLastError();
}

cudaError_t sre_launch_pike_ctx_init(const sre_dev_pike_t &pk, uint8_t *ctx, cudaStream_t stream,
    int *launches)
{
    if (launches) {
        ++*launches;
    }
    k_pike_ctx_init<<<1, 1, pike_smem_words(pk, 0, 1) * 4, stream>>>(pk, ctx);
    return cudaGetLastError();
}

cudaError_t sre_launch_pike_stream(const sre_dev_pike_t &pk, uint8_t *ctx, const uint8_t *buf,
    size_t len, size_t skip, int eof, int want_pending, int64_t *out, uint32_t ovec_slots, cudaStream_t stream,
    int *launches)
{
    (void) want_pending;
    if (launches) {
        ++*launches;
    }
    /* the marks persist between calls, so they stay in the ctx block */
    size_t smem;
    const int parts = smem_parts(pk, 1, false, &smem);
    k_pike_stream<<<1, 1, smem, stream>>>(pk, ctx, buf, len, skip, eof, out, ovec_slots, parts);
    return cudaGetLastError();
}

/* batched streaming Pike: nstreams persistent contexts of pk.ctx_stride bytes each */
cudaError_t sre_launch_pike_streams_init(const sre_dev_pike_t &pk, uint8_t *ctxs, size_t nstreams,
    cudaStream_t stream, int *launches)
{
    if (nstreams == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    k_pike_streams_init<<<(unsigned) ((nstreams + 63) / 64), 64, pike_smem_words(pk, 0, 64) * 4, stream>>>(pk, ctxs,
                                                                                                          nstreams);
    return cudaGetLastError();
}

cudaError_t sre_launch_pike_streams(const sre_dev_pike_t &pk, uint8_t *ctxs, size_t nstreams, const uint8_t *buf,
    const int64_t *offsets, const uint8_t *eofs, int eof_all, int64_t *out, uint32_t ovec_slots,
    cudaStream_t stream, int *launches)
{
    if (nstreams == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    static bool opted = false;
    if (!opted) {
        opted = true;
        cudaFuncSetAttribute(k_pike_streams, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    }
    /* the marks persist between calls, so they stay in the context blocks */
    size_t smem;
    const int parts = smem_parts(pk, 64, false, &smem);
    k_pike_streams<<<(unsigned) ((nstreams + 63) / 64), 64, smem, stream>>>(pk, ctxs, nstreams, buf, offsets, eofs,
                                                                           eof_all, out, ovec_slots, parts);
    return cudaGetLastError();
}
