/*
 * sre_pike_small.cu -- the Pike VM for small single-regex programs with the
 * whole per-line context in shared memory and registers.
 *
 * Same algorithm and the same reference citations as sre_pike.cu (it is the
 * batch-only, eof=1, fresh-context specialisation of pike_exec there:
 * sre_vm_pike_exec sre_vm_pike.c:148-689, add_thread :756-942), for programs
 * with <= 64 instructions, one regex and <= M capture slots:
 *
 *   - dedup tags become two 64-bit register masks: "tag == ctx->tag" and
 *     "tag == ctx->tag - 1" (the only two values the reference ever compares
 *     against within a step; the assertion_hold path temporarily decrements
 *     the tag, :506-528).  ctx->tag++ is `prev = cur; cur = 0`.
 *   - clist is a read-only array plus a small LIFO for the look-ahead closures
 *     the reference prepends to it; nlist is an append-only array.
 *   - thread records (pc, seen_word, M capture slots as int32 line offsets),
 *     the working capture and the DFS stack live in shared memory, laid out
 *     [element][thread] so that every lane only ever touches its own bank.
 *
 * Any line that needs more than the fixed capacities is reported as
 * SRE_K_RETRY and re-run by the general kernel (k_pike_lines), so results are
 * always those of the reference algorithm.  The first-byte prefilter
 * (:256-309) is omitted: it never changes a result, and the start hint computed
 * by k_dfa_lines_hint already skips further.
 */
#include "sre_kernels.cuh"

namespace {

enum {
    OP_CHAR = 1, OP_MATCH = 2, OP_JMP = 3, OP_SPLIT = 4, OP_ANY = 5, OP_SAVE = 6,
    OP_IN = 7, OP_NOTIN = 8, OP_ASSERT = 9
};
enum { AS_SMALL_Z = 0x01, AS_DOLLAR = 0x02, AS_BIG_B = 0x04, AS_SMALL_B = 0x08,
       AS_BIG_A = 0x10, AS_CARET = 0x20 };

constexpr int BLOCK = 64;     /* threads per block; the context size decides blocks per SM */

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

__device__ __forceinline__ sre_dev_inst_t load_inst(const sre_dev_pike_t &pk, int32_t pc)
{
    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(pk.insts + pc));
    sre_dev_inst_t in;
    in.opcode = (uint8_t) (q.x & 0xff);
    in.ch = (uint8_t) ((q.x >> 8) & 0xff);
    in.nranges = (uint16_t) (q.x >> 16);
    in.x = (int32_t) q.y;
    in.y = (int32_t) q.z;
    in.v = (int32_t) q.w;
    return in;
}

/*
 * K: threads per list, M: capture slots, H: hold-stack records, DS: DFS stack.
 * C16: capture values are 16-bit line offsets, two per word (lines shorter than
 * 32 KB): half the shared memory per context, twice the resident warps.
 */
template <int K, int M, int H, int DS, bool C16>
struct small_ctx_t {
    /* words per capture vector */
    static constexpr int MW = C16 ? (M + 1) / 2 : M;
    /* shared-memory sections, in words per lane */
    static constexpr int CAP = 0;
    static constexpr int MAT = CAP + MW;
    static constexpr int L0PC = MAT + MW;
    static constexpr int L0CAP = L0PC + K;
    static constexpr int L1PC = L0CAP + K * MW;
    static constexpr int L1CAP = L1PC + K;
    static constexpr int HSPC = L1CAP + K * MW;
    static constexpr int HSCAP = HSPC + H;
    static constexpr int DSA = HSCAP + H * MW;
    static constexpr int DSB = DSA + DS;
    static constexpr int WORDS = DSB + (C16 ? 0 : DS);

    int32_t    *sm;         /* base + threadIdx.x; stride BLOCK */
    uint64_t    m_cur, m_prev;
    uint32_t    nslots;
    int32_t     matched_id;
    bool        overflow;

    __device__ __forceinline__ int32_t &w(int e) { return sm[e * BLOCK]; }

    /* capture slot `slot` of the vector starting at word section `sec` */
    __device__ __forceinline__ int32_t cap_get(int sec, uint32_t slot)
    {
        if (!C16) {
            return w(sec + slot);
        }
        const int32_t word = w(sec + (slot >> 1));
        return (slot & 1) ? (word >> 16) : (int32_t) (int16_t) (word & 0xffff);
    }
    __device__ __forceinline__ void cap_set(int sec, uint32_t slot, int32_t v)
    {
        if (!C16) {
            w(sec + slot) = v;
            return;
        }
        int32_t &word = w(sec + (slot >> 1));
        word = (slot & 1) ? ((word & 0xffff) | (v << 16)) : ((word & ~0xffff) | (v & 0xffff));
    }
    /* number of words holding nslots captures */
    __device__ __forceinline__ uint32_t cap_words() const { return C16 ? (nslots + 1) >> 1 : nslots; }

    /* tag test / set against ctx->tag (hold == false) or ctx->tag - 1 */
    __device__ __forceinline__ bool tagged(int32_t pc, bool hold) const
    {
        return ((hold ? m_prev : m_cur) >> pc) & 1;
    }
    __device__ __forceinline__ void tag(int32_t pc, bool hold)
    {
        const uint64_t bit = 1ull << pc;
        if (hold) {
            m_prev |= bit;
            m_cur &= ~bit;
        } else {
            m_cur |= bit;
            m_prev &= ~bit;
        }
    }

    /* DFS stack entry: visit pc (slot < 0) or restore capture slot to a value */
    __device__ __forceinline__ void ds_push(int &sp, int32_t a, int32_t slot)
    {
        if (C16) {
            w(DSA + sp) = slot < 0 ? a : (int32_t) (0x80000000u | ((uint32_t) slot << 16) | ((uint32_t) a & 0xffff));
        } else {
            w(DSA + sp) = a;
            w(DSB + sp) = slot;
        }
        sp++;
    }
    __device__ __forceinline__ void ds_pop(int sp, int32_t &a, int32_t &slot)
    {
        if (C16) {
            const int32_t e = w(DSA + sp);
            if (e < 0) {
                slot = (e >> 16) & 0x7fff;
                a = (int32_t) (int16_t) (e & 0xffff);
            } else {
                slot = -1;
                a = e;
            }
        } else {
            a = w(DSA + sp);
            slot = w(DSB + sp);
        }
    }

    /*
     * add_thread (sre_vm_pike.c:756-942).  Appends thread records to the array
     * at (pc_sec, cap_sec) with capacity `cap_n`, counting in *n.  The working
     * capture is w(CAP..).  Returns 0 ok, 1 done (MATCH reached with
     * want_done), -1 overflow.
     */
    __device__ int add_thread(const sre_dev_pike_t &pk, int pc_sec, int cap_sec, int cap_n, int *n,
                              int32_t pc0, int32_t pos, const uint8_t *buffer, bool want_done, bool hold)
    {
        int sp = 0;
        ds_push(sp, pc0, -1);
        const uint32_t ncw = cap_words();
        while (sp > 0) {
            sp--;
            int32_t a, b;
            ds_pop(sp, a, b);
            if (b >= 0) {               /* restore slot b to value a */
                cap_set(CAP, b, a);
                continue;
            }
            int32_t pc = a;
            for (;;) {
                const sre_dev_inst_t in = load_inst(pk, pc);
                uint32_t seen_word = 0;
                bool add = false;

                if (tagged(pc, hold)) {
                    if (in.opcode == OP_SPLIT && !tagged(in.y, hold)) {     /* :770-786 */
                        pc = in.y;
                        continue;
                    }
                    break;
                }
                tag(pc, hold);

                switch (in.opcode) {
                case OP_JMP:
                    pc = in.x;
                    continue;
                case OP_SPLIT:
                    if (sp >= DS) {
                        return -1;
                    }
                    ds_push(sp, in.y, -1);
                    pc = in.x;
                    continue;
                case OP_SAVE:
                    if (sp >= DS) {
                        return -1;
                    }
                    ds_push(sp, cap_get(CAP, in.v), in.v);
                    cap_set(CAP, in.v, pos);
                    pc++;
                    continue;
                case OP_ASSERT:
                    switch (in.v) {
                    case AS_BIG_A:
                        if (pos) {
                            break;
                        }
                        pc++;
                        continue;
                    case AS_CARET:
                        if (pos != 0 && buffer[pos - 1] != '\n') {
                            break;
                        }
                        pc++;
                        continue;
                    case AS_SMALL_B:
                    case AS_BIG_B:
                        seen_word = pos == 0 ? 0 : isword(buffer[pos - 1]);
                        add = true;
                        break;
                    default:
                        add = true;
                        break;
                    }
                    break;
                case OP_MATCH:
                    if (want_done) {
                        for (uint32_t i = 0; i < ncw; i++) {
                            w(MAT + i) = w(CAP + i);
                        }
                        matched_id = in.v;
                        return 1;
                    }
                    add = true;
                    break;
                default:
                    add = true;
                    break;
                }

                if (add) {
                    if (*n >= cap_n) {
                        return -1;
                    }
                    const int t = *n;
                    w(pc_sec + t) = pc | (int32_t) (seen_word << 16);
                    for (uint32_t i = 0; i < ncw; i++) {
                        w(cap_sec + t * MW + i) = w(CAP + i);
                    }
                    *n = t + 1;
                }
                break;
            }
        }
        return 0;
    }
};

template <int K, int M, int H, int DS, bool C16>
__global__ void __launch_bounds__(BLOCK)
k_pike_small(sre_dev_pike_t pk, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
             size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines,
             const int32_t *__restrict__ start_hint, int32_t *__restrict__ rc, int64_t *__restrict__ ovec,
             uint32_t ovec_slots)
{
    extern __shared__ int32_t smem_words[];
    typedef small_ctx_t<K, M, H, DS, C16> ctx_t;
    constexpr int MW = ctx_t::MW;
    ctx_t c;
    c.sm = smem_words + threadIdx.x;
    c.nslots = pk.nslots;

    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    const size_t nwork = lines.list ? (size_t) *lines.count : nlines;
    for (size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < nwork; k += nthreads) {
        const size_t line = lines.list ? (size_t) lines.list[k] : k;
        int64_t *ov = ovec + line * ovec_slots;
        const size_t start = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t end = offsets ? (size_t) offsets[line + 1] : start + linelen;
        const uint8_t *input = buf + start;
        const int32_t size = (int32_t) (end - start);
        int32_t sp = start_hint ? start_hint[line] : 0;

        c.m_cur = c.m_prev = 0;
        c.overflow = false;
        c.matched_id = 0;
        bool matched = false;
        int cur = 0, ncl = 0, nnl = 0, hs = 0;

        const uint32_t ncw = c.cap_words();
        for (uint32_t i = 0; i < ncw; i++) {
            c.w(ctx_t::CAP + i) = -1;
        }
        /* first_buf: the initial closure at the start offset, :202-216 */
        if (c.add_thread(pk, ctx_t::L0PC, ctx_t::L0CAP, K, &ncl, 0, sp, input, false, false) < 0) {
            c.overflow = true;
        }

        for (; !c.overflow && sp <= size; sp++) {
            if (ncl == 0) {
                break;
            }
            c.m_prev = c.m_cur;         /* ctx->tag++ */
            c.m_cur = 0;
            const bool at_end = (sp == size);
            const uint32_t byte = at_end ? 0 : input[sp];
            const bool cur_word = !at_end && isword(byte);
            const int cl_pc = cur ? ctx_t::L1PC : ctx_t::L0PC, cl_cap = cur ? ctx_t::L1CAP : ctx_t::L0CAP;
            const int nl_pc = cur ? ctx_t::L0PC : ctx_t::L1PC, nl_cap = cur ? ctx_t::L0CAP : ctx_t::L1CAP;
            int i = 0;

            for (;;) {
                /* next thread in priority order: pending look-ahead closures first */
                int tp, tc;
                if (hs > 0) {
                    hs--;
                    tp = ctx_t::HSPC + hs;
                    tc = ctx_t::HSCAP + hs * MW;
                } else if (i < ncl) {
                    tp = cl_pc + i;
                    tc = cl_cap + i * MW;
                    i++;
                } else {
                    break;
                }
                const int32_t rec = c.w(tp);
                const int32_t pc = rec & 0xffff;
                const bool t_sw = (rec >> 16) & 1;
                const sre_dev_inst_t in = load_inst(pk, pc);
                bool got_match = false;

                if (in.opcode == OP_ASSERT) {               /* :449-528 */
                    bool hold = false;
                    switch (in.v) {
                    case AS_SMALL_Z: hold = at_end; break;
                    case AS_DOLLAR:  hold = at_end || byte == '\n'; break;
                    case AS_BIG_B:   hold = (t_sw == cur_word); break;
                    case AS_SMALL_B: hold = (t_sw != cur_word); break;
                    default: break;
                    }
                    if (hold) {
                        for (uint32_t k = 0; k < ncw; k++) {
                            c.w(ctx_t::CAP + k) = c.w(tc + k);
                        }
                        /* closure with tag - 1, prepended to clist: append it
                         * above the LIFO top, then reverse that segment */
                        int top = hs;
                        if (c.add_thread(pk, ctx_t::HSPC, ctx_t::HSCAP, H, &top, pc + 1, sp, input, false, true)
                            < 0)
                        {
                            c.overflow = true;
                            break;
                        }
                        for (int lo = hs, hi = top - 1; lo < hi; lo++, hi--) {
                            int32_t tmp = c.w(ctx_t::HSPC + lo);
                            c.w(ctx_t::HSPC + lo) = c.w(ctx_t::HSPC + hi);
                            c.w(ctx_t::HSPC + hi) = tmp;
                            for (uint32_t k = 0; k < ncw; k++) {
                                tmp = c.w(ctx_t::HSCAP + lo * MW + k);
                                c.w(ctx_t::HSCAP + lo * MW + k) = c.w(ctx_t::HSCAP + hi * MW + k);
                                c.w(ctx_t::HSCAP + hi * MW + k) = tmp;
                            }
                        }
                        hs = top;
                    }
                } else if (in.opcode == OP_MATCH) {         /* :530-553 */
                    for (uint32_t k = 0; k < ncw; k++) {
                        c.w(ctx_t::MAT + k) = c.w(tc + k);
                    }
                    c.matched_id = in.v;
                    got_match = true;
                } else if (!at_end && consumes(pk, in, byte)) {
                    for (uint32_t k = 0; k < ncw; k++) {
                        c.w(ctx_t::CAP + k) = c.w(tc + k);
                    }
                    const int r = c.add_thread(pk, nl_pc, nl_cap, K, &nnl, pc + 1, sp + 1, input, true, false);
                    if (r < 0) {
                        c.overflow = true;
                        break;
                    }
                    got_match = (r == 1);
                }

                if (got_match) {        /* cut every lower-priority thread */
                    matched = true;
                    hs = 0;
                    break;
                }
            }

            cur ^= 1;                   /* step_done: swap lists */
            ncl = nnl;
            nnl = 0;
            hs = 0;
            if (at_end) {
                break;
            }
        }

        if (c.overflow) {
            rc[line] = SRE_K_RETRY;
            continue;
        }
        if (matched) {
            rc[line] = c.matched_id;
            for (uint32_t i = 0; i < ovec_slots; i++) {
                ov[i] = i < c.nslots ? (int64_t) c.cap_get(ctx_t::MAT, i) : -1;
            }
        } else {
            rc[line] = SRE_K_DECLINED;
            for (uint32_t i = 0; i < ovec_slots; i++) {
                ov[i] = -1;
            }
        }
    }
}

}  // namespace

bool sre_pike_small_applicable(const sre_dev_pike_t &pk)
{
    return pk.nregexes == 1 && pk.len <= 64 && pk.nslots <= 16;
}

cudaError_t sre_launch_pike_small(const sre_dev_pike_t &pk, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines, const int32_t *start, int32_t *rc,
    int64_t *ovec, uint32_t ovec_slots, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t grid = (nlines + BLOCK - 1) / BLOCK;

#define SRE_SMALL_C(KK, MM, HH, DD, CC)                                                             \
    do {                                                                                            \
        auto kern = k_pike_small<KK, MM, HH, DD, CC>;                                               \
        const size_t smem = (size_t) small_ctx_t<KK, MM, HH, DD, CC>::WORDS * BLOCK * 4;            \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
        if (e != cudaSuccess) return e;                                                             \
        size_t per_sm = (227 * 1024) / (smem + 1024);                                               \
        if (per_sm > 32) per_sm = 32;                                                               \
        const size_t cap = (size_t) sms * (per_sm ? per_sm : 1);                                    \
        if (grid > cap) grid = cap;                                                                 \
        kern<<<(unsigned) grid, BLOCK, smem, stream>>>(pk, buf, offsets, nlines, pitch, linelen, lines,  \
                                                        start, rc, ovec, ovec_slots);               \
    } while (0)
#define SRE_SMALL(KK, MM, HH, DD)                                                                   \
    do {                                                                                            \
        if (c16) SRE_SMALL_C(KK, MM, HH, DD, true); else SRE_SMALL_C(KK, MM, HH, DD, false);        \
    } while (0)

    /* 16-bit capture offsets when every line is shorter than 32 KB */
    const bool c16 = offsets == nullptr && linelen < 32767;
    /* K = 8 threads per list covers 96 % of the reference's t/ corpus (longer
     * lists fall back to k_pike_lines); M = capture slots */
    if (pk.nslots <= 4) {
        SRE_SMALL(8, 4, 4, 12);
    } else if (pk.nslots <= 8) {
        SRE_SMALL(8, 8, 4, 12);
    } else if (pk.nslots <= 10) {
        SRE_SMALL(8, 10, 4, 12);
    } else if (pk.nslots <= 12) {
        SRE_SMALL(8, 12, 4, 12);
    } else {
        SRE_SMALL(8, 16, 4, 12);
    }
#undef SRE_SMALL_C
#undef SRE_SMALL
    return cudaGetLastError();
}
