/*
 * sre_pike_lineage.cu -- Pike VM with submatch captures over the determinised
 * thread lists (../lower/sre_pdfa.h), one CUDA thread per line.
 *
 * What it replaces: sre_vm_pike_exec + sre_vm_pike_add_thread + the capture
 * copies (reference sre_vm_pike.c:148-689, :756-942, sre_capture.c), assertions
 * included (look-behind ones pick the start list, look-ahead ones are threads
 * parked in the lists).  Instead of simulating the thread list (k_pike_table:
 * ~500 instructions per byte, 10 of 32 lanes busy), a lane runs
 *
 *   forward   one look-up per byte in the P-DFA (state = the ordered thread
 *             list), storing the state of each position in a 64-entry ring in
 *             shared memory and noting the last step that reported a match;
 *             it ends when the list is empty (every thread of higher priority
 *             than the match has died) or the line does;
 *   backward  from the thread that matched, parent by parent through the
 *             transitions' provenance records, giving each capture slot the
 *             position of the last step that SAVEd it -- until the lineage
 *             reaches the ".*?" thread or the start closure.
 *
 * Every lane executes the same two loops: no per-lane thread lists, no
 * divergence beyond the span lengths.  When the walk back needs a position
 * older than the ring (a match longer than ~60 bytes), the lane runs the
 * automaton forward again up to that position -- the P-DFA is deterministic,
 * the ring then holds the 64 transitions in front of it -- and walks on: the
 * first such pass starts where the search started and notes the state at every
 * 64th position, the later ones start from those; after REFILLS passes (a
 * lineage of ~4 KB) the line is reported SRE_K_RETRY and re-run by k_pike_table.
 */
#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

constexpr int RING = 64;            /* positions remembered per lane (power of two) */
constexpr int REFILLS = 64;         /* forward re-runs per line before it is handed to the next tier */
constexpr int CKPTS = 32;           /* states remembered by the first re-run, one per RING positions */

/* bytes of the tables a block keeps in shared memory next to the rings */
__host__ __device__ inline size_t lineage_table_bytes(const sre_dev_pdfa_t &d)
{
    const size_t T = (size_t) d.nstates * d.nclasses;
    return align_up(T * 2, 16) + align_up((T + 1) * 4, 16) + (size_t) d.nent * 8 + T * 8
           + align_up((size_t) d.nstates * 4, 16);
}

/*
 * MODE 0: every table in shared memory and the transitions as a [state][byte]
 * u32 table { next | match flag, transition number << 16 } (small automata:
 * C3's has 22 lists) -- per input byte one PRMT (the byte spliced under the
 * state), one LDS and one STS (the ring);
 * MODE 1: class-indexed tables in shared memory; MODE 2: class-indexed tables
 * through the read-only path (large automata).  RT: the ring's element type.
 * The ring holds the transition taken at each position (state * classes +
 * class; one byte per position when the automaton has at most 256
 * transitions) and is indexed by the low bits of the byte's ADDRESS, so that
 * the 16 positions of an aligned load are 16 consecutive ring slots.
 * prefilled: the ovector rows are -1 already.
 */
template <int MODE, typename RT, int TB>
__global__ void __launch_bounds__(TB, TB == 1024 ? 2 : 1)
k_pike_lineage(sre_dev_pdfa_t d, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
               size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines,
               const int32_t *__restrict__ start_hint, int32_t *__restrict__ rc, int64_t *__restrict__ ovec,
               uint32_t ovec_slots, int prefilled, sre_pike_work_t *work)
{
    constexpr bool SMEM_TAB = MODE != 2, BYTE = MODE == 0;
    extern __shared__ __align__(16) uint8_t smem[];
    /* [rings][clsmap 256][trans | eofs | ent | mev | eof (SMEM_TAB)][t256 (BYTE)].  A warp's rings
     * are interleaved word by word, so that lane l only ever touches bank l: slot i of lane l is
     * element i % EPW of word (i / EPW) * 32 + l of the warp's region (no bank conflicts,
     * whatever positions the lanes are at) */
    constexpr int EPW = 4 / (int) sizeof(RT);
    uint8_t *ring = smem + (threadIdx.x >> 5) * (RING * 32 * sizeof(RT)) + (threadIdx.x & 31) * 4;
    auto ring_at = [&](uint32_t i) -> RT * {
        return reinterpret_cast<RT *>(ring + (i / EPW) * 128 + (i % EPW) * sizeof(RT));
    };
    uint8_t *s_cls = smem + RING * TB * sizeof(RT);
    const uint16_t *trans = d.trans;
    const uint32_t *t256 = nullptr;
    const uint32_t *eofs = d.eofs, *eof = d.eof;
    const uint2 *ent = d.ent, *mev = d.mev;
    const uint32_t C = d.nclasses, T = d.nstates * C;
    for (uint32_t i = threadIdx.x; i < 256; i += TB) {
        s_cls[i] = d.clsmap[i];
    }
    if (SMEM_TAB) {
        uint8_t *p = s_cls + 256;
        uint16_t *s_trans = reinterpret_cast<uint16_t *>(p);
        p += align_up((size_t) T * 2, 16);
        uint32_t *s_eofs = reinterpret_cast<uint32_t *>(p);
        p += align_up((size_t) (T + 1) * 4, 16);
        uint2 *s_ent = reinterpret_cast<uint2 *>(p);
        p += (size_t) d.nent * 8;
        uint2 *s_mev = reinterpret_cast<uint2 *>(p);
        p += (size_t) T * 8;
        uint32_t *s_eof = reinterpret_cast<uint32_t *>(p);
        p += align_up((size_t) d.nstates * 4, 16);
        for (uint32_t i = threadIdx.x; i < T; i += TB) {
            s_trans[i] = d.trans[i];
            s_mev[i] = d.mev[i];
        }
        for (uint32_t i = threadIdx.x; i <= T; i += TB) {
            s_eofs[i] = d.eofs[i];
        }
        for (uint32_t i = threadIdx.x; i < d.nent; i += TB) {
            s_ent[i] = d.ent[i];
        }
        for (uint32_t i = threadIdx.x; i < d.nstates; i += TB) {
            s_eof[i] = d.eof[i];
        }
        trans = s_trans;
        eofs = s_eofs;
        ent = s_ent;
        mev = s_mev;
        eof = s_eof;
        if (BYTE) {
            uint32_t *s_t256 = reinterpret_cast<uint32_t *>(p);
            for (uint32_t i = threadIdx.x; i < d.nstates * 256; i += TB) {
                const uint32_t t = (i >> 8) * C + d.clsmap[i & 0xff];
                s_t256[i] = d.trans[t] | (t << 16);
            }
            t256 = s_t256;
        }
    }
    __syncthreads();
    auto ld16 = [&](const uint16_t *q) -> uint32_t { return SMEM_TAB ? *q : __ldg(q); };
    auto ld32 = [&](const uint32_t *q) -> uint32_t { return SMEM_TAB ? *q : __ldg(q); };
    auto ld64 = [&](const uint2 *q) -> uint2 { return SMEM_TAB ? *q : __ldg(q); };
    const size_t nwork = lines.list ? (size_t) *lines.count : nlines;

    for (size_t k = (size_t) blockIdx.x * TB + threadIdx.x; k < nwork; k += (size_t) gridDim.x * TB) {
        const size_t line = lines.list ? (size_t) lines.list[k] : k;
        const size_t lstart = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t lend = offsets ? (size_t) offsets[line + 1] : lstart + linelen;
        const uint8_t *input = buf + lstart;
        const int32_t size = (int32_t) (lend - lstart);
        int64_t *ov = ovec + line * ovec_slots;
        const uint32_t in_lo = (uint32_t) reinterpret_cast<uintptr_t>(input);
        /* The search may begin at ANY offset up to the hint: the hint is a position at which no
         * thread is alive, so the threads a start before it adds are dead again when it is
         * reached.  Beginning at the 16-byte boundary below it makes every block a whole one. */
        int32_t start = start_hint ? start_hint[line] : 0;
        {
            const int32_t down = (int32_t) ((in_lo + (uint32_t) start) & 15u);
            start = start - down < 0 ? 0 : start - down;
        }
        /* the ring slot of position p */
        auto slot = [&](int32_t p) -> RT * { return ring_at((in_lo + (uint32_t) p) & (RING - 1)); };

        /* forward: the transition taken at every position, the last one that reported a match */
        /* the start list by what lies before the first byte looked at (the look-behind context of
         * `\A` / `^`): nothing, a newline, anything else */
        uint32_t v0 = 0;
        if (d.ctx_dep && start > 0) {
            const uint32_t pb = __ldg(input + start - 1);
            const bool word = (pb - '0' < 10u) || ((pb | 0x20) - 'a' < 26u) || pb == '_';
            v0 = pb == '\n' ? 1u : word ? 2u : 3u;
        }
        uint32_t s = d.init[v0];
        int32_t pos = start, mpos = -1;
        auto step = [&](uint32_t b, int32_t p, RT *where) {
            uint32_t e;
            if (BYTE) {
                e = t256[(s << 8) | b];
                *where = (RT) (e >> 16);
            } else {
                const uint32_t t = s * C + s_cls[b];
                *where = (RT) t;
                e = ld16(trans + t);
            }
            mpos = (e & 0x8000u) ? p : mpos;
            s = e & 0x7fffu;
        };
        /* the aligned 16 bytes that hold input[pos]; the block after it is requested before
         * this one is processed */
        uint4 vnext = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<uintptr_t>(input + pos) & ~(uintptr_t) 15));
        while (pos < size && s != 0) {
            /* byte q of the block is position rel + q */
            const uintptr_t at = reinterpret_cast<uintptr_t>(input + pos);
            const uint4 v = vnext;
            const int32_t rel = pos - (int32_t) (at & 15);
            if (rel + 16 < size) {
                vnext = __ldg(reinterpret_cast<const uint4 *>((at & ~(uintptr_t) 15) + 16));
            }
            /* 16 consecutive slots (an aligned block never wraps): slot q of them at a fixed offset */
            uint8_t *rb = reinterpret_cast<uint8_t *>(slot(rel));
            auto blk = [&](int q) -> RT * {
                return reinterpret_cast<RT *>(rb + (q / EPW) * 128 + (q % EPW) * (int) sizeof(RT));
            };
            if (rel == pos && rel + 16 <= size) {
                /* a whole block: the liveness test once per 4 bytes (the empty list stays
                 * empty, so the few transitions taken after it has emptied change nothing) */
                const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    if (s != 0) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            step((w[g] >> (8 * i)) & 0xff, rel + 4 * g + i, blk(4 * g + i));
                        }
                        pos = rel + 4 * g + 4;
                    }
                }
            } else {
                /* the first block of a line that begins off a 16-byte boundary, or the last one */
                const uint32_t w[4] = { v.x, v.y, v.z, v.w };
                const int32_t q0 = pos - rel, q1 = size - rel < 16 ? size - rel : 16;
#pragma unroll 1
                for (int32_t q = q0; q < q1 && s != 0; q++) {
                    const uint32_t word = q < 8 ? (q < 4 ? w[0] : w[1]) : (q < 12 ? w[2] : w[3]);
                    step((word >> ((q & 3) * 8)) & 0xff, rel + q, slot(rel + q));
                    pos = rel + q + 1;
                }
            }
        }
        const uint32_t ef = s != 0 ? ld32(eof + s) : 0xffu;
        const bool at_eof = (ef & 0xff) != 0xff;
        if (mpos < 0 && !at_eof) {
            rc[line] = SRE_K_DECLINED;
            if (!prefilled) {
                for (uint32_t i = 0; i < ovec_slots; i++) {
                    ov[i] = -1;
                }
            }
            continue;
        }

        /* backward along the lineage of the thread that matched: j = its index in the list
         * before step u + 1; a slot keeps the position of the LAST step that SAVEd it */
        uint32_t j, rid, unset = 0xffffffffu;
        int32_t u;
        bool lost = false, stop = false;
        if (!prefilled) {
            for (uint32_t i = 0; i < ovec_slots; i++) {
                ov[i] = -1;
            }
        }
        auto assign = [&](uint32_t mask, int32_t where) {
            uint32_t m = mask & unset;
            unset &= ~mask;
            while (m) {
                const uint32_t sl = __ffs(m) - 1;
                m &= m - 1;
                if (sl < ovec_slots) {
                    ov[sl] = where;
                }
            }
        };
        /* the transition taken at position p, as an index into eofs / mev */
        auto taken = [&](int32_t p) -> uint32_t { return *slot(p); };
        /* positions the ring has lost: forward again from the start of the search, up to and
         * including position `upto` (same automaton, same bytes: the same transitions) */
        int refills = 0;
        /* the first re-run starts at the start of the search and notes the state in front of every
         * RING-th position; the later ones start from the last such state that still lets them
         * cover the RING positions up to `upto` (<= 2 * RING - 1 steps) */
        uint16_t ckpt[CKPTS];
        int32_t nck = 0;
        auto refill = [&](int32_t upto) -> bool {
            if (refills == REFILLS) {
                return false;
            }
            uint32_t r = d.init[v0];
            int32_t p = start;
            if (refills > 0) {
                int32_t i = (upto - (RING - 1) - start) / RING;
                i = upto - (RING - 1) < start ? 0 : i >= nck ? nck - 1 : i;
                r = ckpt[i];
                p = start + i * RING;
            }
            const bool note = refills == 0;
            refills++;
            for (; p <= upto; p++) {
                if (note && ((p - start) & (RING - 1)) == 0 && nck < CKPTS) {
                    ckpt[nck++] = (uint16_t) r;
                }
                const uint32_t b = __ldg(input + p);
                uint32_t e;
                if (BYTE) {
                    e = t256[(r << 8) | b];
                    *slot(p) = (RT) (e >> 16);
                } else {
                    const uint32_t t = r * C + s_cls[b];
                    *slot(p) = (RT) t;
                    e = ld16(trans + t);
                }
                r = e & 0x7fffu;
            }
            return true;
        };
        int32_t oldest = pos - RING;            /* positions > oldest are in the ring */
        if (at_eof) {
            if (d.eof0 != nullptr) {
                assign(__ldg(d.eof0 + s), size);        /* SAVEd by assertions resolved at the end */
            }
            j = ef & 0xff;
            rid = ef >> 16;
            u = size - 1;
        } else if (mpos <= oldest && !refill(mpos)) {
            lost = true;
            j = rid = 0;
            u = -1;
        } else {
            if (refills) {
                oldest = mpos - RING;
            }
            const uint32_t tm = taken(mpos);
            const uint2 m = ld64(mev + tm);
            assign(m.x, mpos + 1);
            if (d.mev0 != nullptr) {
                assign(__ldg(d.mev0 + tm), mpos);       /* ... by assertions resolved before the byte */
            }
            j = m.y & 0xff;
            stop = (m.y & 0x100u) != 0;
            rid = m.y >> 16;
            u = mpos - 1;
        }
        while (!lost && !stop) {
            if (u < start) {
                if (j != d.init_any[v0]) {
                    assign(__ldg(d.init_mask + d.init_mask_ofs[v0] + j), start);   /* a thread of the start closure */
                }
                break;
            }
            if (u <= oldest) {
                if (!refill(u)) {
                    lost = true;
                    break;
                }
                oldest = u - RING;
            }
            const uint32_t ei = ld32(eofs + taken(u)) + j;
            const uint2 e = ld64(ent + ei);
            assign(e.x, u + 1);
            if (d.ent0 != nullptr) {
                assign(__ldg(d.ent0 + ei), u);
            }
            j = e.y & 0xff;
            stop = (e.y & 0x100u) != 0;         /* the parent is the ".*?" thread: it carries no captures */
            u--;
        }
        if (lost) {
            /* the lineage is older than the ring: the next tier re-runs the line */
            rc[line] = SRE_K_RETRY;
            atomicAdd(&work->given_up[0], 1u);
            continue;
        }
        rc[line] = (int32_t) rid;
    }
}

}  // namespace

bool sre_pike_lineage_applicable(const sre_dev_pdfa_t &d, size_t linelen)
{
    return d.nstates != 0 && linelen < (1ull << 31);
}

/* same contract as sre_launch_pike_table, pass 0: lines it gives up on get SRE_K_RETRY and are
 * counted in work->given_up[0]; prefilled: the ovector rows are -1 already */
cudaError_t sre_launch_pike_lineage(const sre_dev_pdfa_t &d, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines, const int32_t *start, int32_t *rc,
    int64_t *ovec, uint32_t ovec_slots, int prefilled, sre_pike_work_t *work, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    cudaError_t ce = cudaMemsetAsync(work, 0, sizeof(*work), stream);
    if (ce != cudaSuccess) {
        return ce;
    }
    const size_t T = (size_t) d.nstates * d.nclasses;
    const size_t tab = lineage_table_bytes(d), tab256 = (size_t) d.nstates * 1024;
    /* 0: small automaton, [state][byte] table and everything else in shared memory, blocks of 1024
     * (two per SM); 1: class tables in shared memory; 2: tables through the read-only path */
    const int mode = (tab + tab256 <= 40 * 1024 && T < 65536) ? 0 : tab <= 40 * 1024 ? 1 : 2;
    const int rt = T <= 256 ? 1 : T <= 65535 ? 2 : 4;
    const int tb = mode == 0 ? 1024 : 128;
    const size_t smem = (size_t) RING * tb * rt + 256 + (mode != 2 ? tab : 0) + (mode == 0 ? tab256 : 0);
    size_t grid = (nlines + tb - 1) / tb;
    size_t per_sm = (220 * 1024) / (smem + 1024), by_threads = 2048 / tb;
    per_sm = per_sm > by_threads ? by_threads : per_sm < 1 ? 1 : per_sm;
    const size_t cap = (size_t) num_sms() * per_sm;
    if (grid > cap) {
        grid = cap;
    }
    typedef void (*kern_t)(sre_dev_pdfa_t, const uint8_t *, const int64_t *, size_t, size_t, size_t, sre_line_list_t,
                           const int32_t *, int32_t *, int64_t *, uint32_t, int, sre_pike_work_t *);
    static const kern_t kerns[3][3] = {
        { k_pike_lineage<0, uint8_t, 1024>, k_pike_lineage<0, uint16_t, 1024>, nullptr },
        { k_pike_lineage<1, uint8_t, 128>, k_pike_lineage<1, uint16_t, 128>, k_pike_lineage<1, uint32_t, 128> },
        { k_pike_lineage<2, uint8_t, 128>, k_pike_lineage<2, uint16_t, 128>, k_pike_lineage<2, uint32_t, 128> },
    };
    const int ri = rt == 1 ? 0 : rt == 2 ? 1 : 2;
    const kern_t kern = kerns[mode][ri];
    static bool opted[3][3];
    if (!opted[mode][ri]) {
        ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (ce != cudaSuccess) {
            return ce;
        }
        opted[mode][ri] = true;
    }
    kern<<<(unsigned) grid, tb, smem, stream>>>(d, buf, offsets, nlines, pitch, linelen, lines, start, rc, ovec,
                                               ovec_slots, prefilled, work);
    return cudaGetLastError();
}
