/*
 * sre_thompson.cu -- Thompson (boolean) match kernels for sm_100a.
 *
 * What they replace: sre_vm_thompson_exec (reference sre_vm_thompson.c:63-270)
 * and the code its x86-64 JIT emits (sre_vm_thompson_x64.dasc:746-886).  The
 * thread-list loop becomes table look-ups over the lowered program
 * (../lower/sre_lower.h):
 *
 *   k_dfa_lines_tma_early  one CUDA thread per line, determinised program.
 *                  Input tiles (32 lines x 128 bytes per warp) are staged
 *                  through shared memory by the TMA unit (128-byte swizzle: the
 *                  per-lane 16-byte reads of "my row" are bank-conflict free).
 *                  Per input byte: one PRMT (the byte's address in row 0, off the
 *                  chain), one IMAD and one LDS.U8 into the [state][byte] table,
 *                  whose rows are padded to 260 bytes (bank conflicts).
 *   k_dfa_lines_skipw  the same with a warp-uniform word skip for automata whose
 *                  start state is left by few byte values.  HBM-bound.
 *   k_dfa_lines_hint[_skip]  verdict + Pike start hint; packs the matching lines
 *                  for the Pike pass.
 *   k_dfa_lines_big  tiled lines, class table through L1 (automata beyond shared
 *                  memory); bound by the L1/shared data pipe.
 *   k_dfa_generic  same automaton, any alignment / ragged offsets, any table
 *                  size, optional state carry (SRE_AGAIN) -- correctness tier.
 *   k_nfa_packed   no DFA, <= 64 lowered states, <= 4 non-shift movers: thread per
 *                  line, the thread set in a 32/64-bit register, one shared-
 *                  memory entry per byte class.
 *   k_nfa64_lines  the same for any number of non-shift movers (loop over bits).
 *   k_nfa_lines    the general tier: one warp per line, the NFA thread set is a
 *                  warp-wide bitmask (lane l owns words l, l+32, ... of it, up
 *                  to 4096 states), successor sets come from a shift for
 *                  "next pc" states and from follow-row ORs for the rest.
 */
#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

/* ---- k_dfa_lines_tma_early --------------------------------------------------- */

template <bool CLS>
struct line_consumer_t {
    step260_t       st256;      /* rows padded to 260 bytes: lanes in different states do not collide
                                   on the bank of the byte they read (ncu: 16 % of the shared-memory
                                   wavefronts were such conflicts and the pipe was 88 % busy) */
    stepcls_t       stcls;
    const uint8_t  *fin;
    uint32_t        start, acc, s;
    size_t          nlines;
    int32_t        *rc;

    __device__ __forceinline__ void begin(size_t) { s = start; }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        if (CLS) {
            s = stcls.word(stcls.word(stcls.word(stcls.word(s, v.x), v.y), v.z), v.w);
        } else {
            s = st256.word(st256.word(st256.word(st256.word(s, v.x), v.y), v.z), v.w);
        }
    }
    __device__ __forceinline__ void byte(uint32_t b) { s = CLS ? stcls.byte(s, b) : st256.byte(s, b); }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t line = group * 32 + (threadIdx.x & 31);
        if (line < nlines) {
            /* ACC: a step saw a live MATCH thread; fin: the EOF step does */
            rc[line] = (s == acc || fin[s]) ? SRE_K_OK : SRE_K_DECLINED;
        }
    }
};

/* TMA staging with early stage release (see tile_pipeline_tma_early) */
template <int STAGES, bool CLS, int THREADS, int BLOCKS>
__global__ void __launch_bounds__(THREADS, BLOCKS)
k_dfa_lines_tma_early(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, size_t nlines,
                      uint32_t linelen, int32_t *__restrict__ rc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, CLS);
    uint8_t *s_tab = smem;
    uint8_t *s_fin = smem + plan.fin_ofs;
    uint8_t *s_cls = smem + plan.cls_ofs;

    if (CLS) {
        load_table(s_tab, reinterpret_cast<const uint8_t *>(dfa.tcls), plan.tab_bytes);
        load_table(s_cls, dfa.clsmap, 256);
    } else {
        load_table260(s_tab, dfa.t256, dfa.nstates);
    }
    load_table(s_fin, dfa.fin, align_up(dfa.nstates, 16));
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    line_consumer_t<CLS> cons;
    cons.st256.tab_s = smem_u32(s_tab);
    cons.stcls.tab = reinterpret_cast<const uint16_t *>(s_tab);
    cons.stcls.cls = s_cls;
    cons.stcls.ncls = dfa.nclasses;
    cons.fin = s_fin;
    cons.start = dfa.start;
    cons.acc = dfa.acc;
    cons.nlines = nlines;
    cons.rc = rc;

    tile_pipeline_tma_early<STAGES>(cons, &tmap, nlines, linelen,
                                    smem + plan.stage_ofs + (size_t) warp * STAGES * 32 * 128,
                                    reinterpret_cast<uint64_t *>(smem + plan.bar_ofs) + warp * MAX_STAGES,
                                    (size_t) blockIdx.x * warps_per_block + warp,
                                    (size_t) gridDim.x * warps_per_block);
}

/* ---- k_dfa_lines_skipw ------------------------------------------------------ */

/*
 * Word-skip flavour of the early-release kernel, for automata whose start
 * state is left by <= 4 byte values: per 4-byte word every lane tests (3 ALU
 * ops per byte value) whether the word holds a leave byte; if NO lane of the
 * warp is inside a partial match and NO lane's word holds a leave byte, the
 * four table look-ups of that word are skipped for the whole warp (a
 * warp-uniform branch, no divergence).  T[start][b] == start for every skipped
 * byte and ACC is absorbing, so the result is identical to k_dfa_lines.  On log
 * text this removes more than half of the LDS.U8 look-ups that bound
 * k_dfa_lines_tma_early (shared-memory pipe 90 % busy, profiles/).
 */
template <int NPAT>
struct skipw_consumer_t {
    step256_t       st256;
    const uint8_t  *fin;
    uint32_t        start, acc, s;
    uint32_t        pat[4];
    size_t          nlines;
    int32_t        *rc;

    __device__ __forceinline__ void begin(size_t) { s = start; }
    __device__ __forceinline__ uint32_t leave(uint32_t w) const
    {
        uint32_t h = 0;
#pragma unroll
        for (int p = 0; p < NPAT; p++) {
            const uint32_t x = w ^ pat[p];
            h |= (x - 0x01010101u) & ~x & 0x80808080u;     /* some byte of x is zero */
        }
        return h;
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        /* one vote per 16 bytes: bit k = "word k must be looked up by some lane".
         * A looked-up word can leave lanes inside a partial match, so every word
         * after the first needed one is looked up as well. */
        uint32_t bits = (leave(v.x) ? 1u : 0u) | (leave(v.y) ? 2u : 0u) | (leave(v.z) ? 4u : 0u)
                      | (leave(v.w) ? 8u : 0u);
        bits = s == acc ? 0u : (s != start ? 15u : bits);
        uint32_t m = __reduce_or_sync(0xffffffffu, bits);
        m |= m << 1;
        m |= m << 2;
        if (m & 1) s = st256.word(s, v.x);
        if (m & 2) s = st256.word(s, v.y);
        if (m & 4) s = st256.word(s, v.z);
        if (m & 8) s = st256.word(s, v.w);
    }
    __device__ __forceinline__ void byte(uint32_t b) { s = st256.byte(s, b); }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t line = group * 32 + (threadIdx.x & 31);
        if (line < nlines) {
            rc[line] = (s == acc || fin[s]) ? SRE_K_OK : SRE_K_DECLINED;
        }
    }
};

template <int NPAT, int THREADS, int BLOCKS>
__global__ void __launch_bounds__(THREADS, BLOCKS)
k_dfa_lines_skipw(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, size_t nlines,
                  uint32_t linelen, uint4 pats, int32_t *__restrict__ rc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, false);
    load_table(smem, dfa.t256, (size_t) dfa.nstates * 256);
    load_table(smem + plan.fin_ofs, dfa.fin, align_up(dfa.nstates, 16));
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    skipw_consumer_t<NPAT> cons;
    cons.st256.tab = smem;
    cons.fin = smem + plan.fin_ofs;
    cons.start = dfa.start;
    cons.acc = dfa.acc;
    cons.pat[0] = pats.x;
    cons.pat[1] = pats.y;
    cons.pat[2] = pats.z;
    cons.pat[3] = pats.w;
    cons.nlines = nlines;
    cons.rc = rc;
    tile_pipeline_tma_early<1>(cons, &tmap, nlines, linelen, smem + plan.stage_ofs + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + plan.bar_ofs) + warp * MAX_STAGES,
                               (size_t) blockIdx.x * warps_per_block + warp, (size_t) gridDim.x * warps_per_block);
}

/* the lines the gate lets through, appended to the packed list the Pike kernels work from (warp-
 * aggregated; the order is irrelevant); the ovector rows of the others are final: all -1.
 * Called by all 32 lanes. */
__device__ __forceinline__ void gate_pack(const sre_gate_pack_t &pk, bool take, bool valid, size_t line)
{
    if (valid && !take && pk.ovec != nullptr) {
        int64_t *row = pk.ovec + line * pk.ovec_slots;
        for (uint32_t i = 0; i < pk.ovec_slots; i++) {
            row[i] = -1;
        }
    }
    const uint32_t m = __ballot_sync(FULL, take);
    if (m == 0) {
        return;
    }
    const uint32_t lane = threadIdx.x & 31, leader = (uint32_t) __ffs((int) m) - 1;
    uint32_t base = 0;
    if (lane == leader) {
        base = atomicAdd(pk.count, (uint32_t) __popc(m));
    }
    base = __shfl_sync(FULL, base, leader);
    if (take) {
        pk.list[base + __popc(m & ((1u << lane) - 1))] = (uint32_t) line;
    }
}

/* ---- k_dfa_lines_hint ------------------------------------------------------ */

/*
 * Same automaton through the [256][256] "restart" table: bit 7 of an entry says
 * that only the ".*?" thread consumed the byte, i.e. every partial match died
 * on it.  hint = offset just after the last such byte seen before the first
 * match: a leftmost-first (Pike) search may start there instead of at 0, which
 * is the reference's first-byte prefilter idea (sre_vm_pike.c:256-309) made
 * exact by the automaton instead of a thread-list comparison.
 */
struct hint_consumer_t {
    uint32_t        tab_s;      /* h256 in shared memory, rows of ROW260 bytes (shared-window address,
                                   low byte 0) */
    const uint8_t  *fin;
    uint32_t        acc, s, pos;
    uint32_t        fw, fpos;   /* restart flags of the last word that had any, offset just after it */
    size_t          nlines;
    int32_t        *rc, *hint;
    sre_gate_pack_t pack;

    __device__ __forceinline__ void begin(size_t) { s = 0; pos = 0; fw = 0; fpos = 0; }
    /*
     * A word at a time.  The rows are padded to 260 bytes (ncu, round 2: with 256-byte rows 42 %
     * of this kernel's shared-memory wavefronts were bank conflicts -- lanes in different states
     * reading bytes of the same 4-byte group -- and the pipe was 91 % busy); the byte's address
     * within row 0 is one PRMT off the state chain (one IMAD + one LDS per byte on it).  The four
     * states are packed into one register (IMADs: the fma pipe is the idle one) and the restart
     * flags tested together; a word that has any is remembered with two predicated moves, and
     * the hint is worked out of it once, at the end of the line.
     */
    __device__ __forceinline__ void word(uint32_t w, uint32_t after)
    {
        const uint32_t b0 = __byte_perm(w, tab_s, 0x7650), b1 = __byte_perm(w, tab_s, 0x7651);
        const uint32_t b2 = __byte_perm(w, tab_s, 0x7652), b3 = __byte_perm(w, tab_s, 0x7653);
        const uint32_t a0 = step260_t::lds_u8(s * ROW260 + b0);
        const uint32_t a1 = step260_t::lds_u8(a0 * ROW260 + b1);
        const uint32_t a2 = step260_t::lds_u8(a1 * ROW260 + b2);
        const uint32_t a3 = step260_t::lds_u8(a2 * ROW260 + b3);
        s = a3;
        const uint32_t m = ((a3 * 256u + a2) * 65536u + (a1 * 256u + a0)) & 0x80808080u;
        fw = m ? m : fw;
        fpos = m ? after : fpos;
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        word(v.x, pos + 4);
        word(v.y, pos + 8);
        word(v.z, pos + 12);
        word(v.w, pos + 16);
        pos += 16;
    }
    __device__ __forceinline__ void byte(uint32_t b)
    {
        s = step260_t::lds_u8(s * ROW260 + tab_s + b);
        pos++;
        if (s & 0x80) {
            fw = 0x80000000u;
            fpos = pos;
        }
    }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t line = group * 32 + (threadIdx.x & 31);
        bool ok = false;
        if (line < nlines) {
            const uint32_t st = s & 0x7f;
            ok = st == acc || fin[st];
            rc[line] = ok ? SRE_K_OK : SRE_K_DECLINED;
            /* just after the last flagged byte of that word */
            hint[line] = fw ? (int32_t) (fpos - ((uint32_t) __clz((int) fw) >> 3)) : 0;
        }
        if (pack.list != nullptr) {
            gate_pack(pack, ok, line < nlines, line);
        }
    }
};

/* shared memory of the hint kernels: [256 rows x 260 B][fin 256][barriers][stages] */
constexpr size_t HINT_FIN_OFS = 256 * ROW260, HINT_BAR_OFS = HINT_FIN_OFS + 256,
                 HINT_STAGE_OFS = (HINT_BAR_OFS + MAX_WARPS * MAX_STAGES * 8 + 1023) / 1024 * 1024;

__global__ void __launch_bounds__(1024, 1)
k_dfa_lines_hint(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, size_t nlines,
                 uint32_t linelen, int32_t *__restrict__ rc, int32_t *__restrict__ hint, sre_gate_pack_t pack)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    load_table260(smem, dfa.h256, 256);
    load_table(smem + HINT_FIN_OFS, dfa.fin, align_up(dfa.nstates, 16));
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    hint_consumer_t cons;
    cons.tab_s = (uint32_t) __cvta_generic_to_shared(smem);
    cons.fin = smem + HINT_FIN_OFS;
    cons.acc = dfa.acc;
    cons.nlines = nlines;
    cons.rc = rc;
    cons.hint = hint;
    cons.pack = pack;
    tile_pipeline_tma_early<1>(cons, &tmap, nlines, linelen,
                               smem + HINT_STAGE_OFS + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + HINT_BAR_OFS) + warp * MAX_STAGES,
                               (size_t) blockIdx.x * warps_per_block + warp,
                               (size_t) gridDim.x * warps_per_block);
}

/*
 * The same with the word skip of k_dfa_lines_skipw (NPAT = 1 or 2 "leave"
 * bytes of the start state): a 4-byte word no lane needs is not looked up.  A
 * lane that skips a word sits in the start state and the word holds no leave
 * byte, so each of its bytes is one "only the .*? thread consumed it" and the
 * hint moves to the end of the word; a lane already in ACC keeps its hint.
 */
template <int NPAT>
struct hint_skip_consumer_t {
    const uint8_t  *tab;        /* h256 in shared memory */
    const uint8_t  *fin;
    uint32_t        acc, start, s, pos, p0;
    uint32_t        pat[2];
    size_t          nlines;
    int32_t        *rc, *hint;
    sre_gate_pack_t pack;

    __device__ __forceinline__ void begin(size_t) { s = start; pos = 0; p0 = 0; }
    __device__ __forceinline__ void step(uint32_t addr)
    {
        s = tab[addr];
        pos++;
        if (s & 0x80) {
            p0 = pos;
        }
    }
    __device__ __forceinline__ uint32_t leave(uint32_t w) const
    {
        uint32_t h = 0;
#pragma unroll
        for (int p = 0; p < NPAT; p++) {
            const uint32_t x = w ^ pat[p];
            h |= (x - 0x01010101u) & ~x & 0x80808080u;
        }
        return h;
    }
    __device__ __forceinline__ void word(uint32_t w, bool look)
    {
        if (look) {
            step(__byte_perm(w, s, 0x5540));
            step(__byte_perm(w, s, 0x5541));
            step(__byte_perm(w, s, 0x5542));
            step(__byte_perm(w, s, 0x5543));
        } else {
            pos += 4;
            if ((s & 0x7f) == start) {
                p0 = pos;
            }
        }
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        const uint32_t st = s & 0x7f;
        uint32_t bits = (leave(v.x) ? 1u : 0u) | (leave(v.y) ? 2u : 0u) | (leave(v.z) ? 4u : 0u)
                      | (leave(v.w) ? 8u : 0u);
        bits = st == acc ? 0u : (st != start ? 15u : bits);
        uint32_t m = __reduce_or_sync(0xffffffffu, bits);
        m |= m << 1;
        m |= m << 2;
        word(v.x, m & 1);
        word(v.y, m & 2);
        word(v.z, m & 4);
        word(v.w, m & 8);
    }
    __device__ __forceinline__ void byte(uint32_t b) { step((s << 8) | b); }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t line = group * 32 + (threadIdx.x & 31);
        bool ok = false;
        if (line < nlines) {
            const uint32_t st = s & 0x7f;
            ok = st == acc || fin[st];
            rc[line] = ok ? SRE_K_OK : SRE_K_DECLINED;
            hint[line] = (int32_t) p0;
        }
        if (pack.list != nullptr) {
            gate_pack(pack, ok, line < nlines, line);
        }
    }
};

template <int NPAT>
__global__ void __launch_bounds__(1024, 1)
k_dfa_lines_hint_skip(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, size_t nlines,
                      uint32_t linelen, uint32_t pat0, uint32_t pat1, int32_t *__restrict__ rc,
                      int32_t *__restrict__ hint, sre_gate_pack_t pack)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const dfa_smem_plan_t plan = dfa_smem_plan(256, 0, false);
    load_table(smem, dfa.h256, 65536);
    load_table(smem + plan.fin_ofs, dfa.fin, align_up(dfa.nstates, 16));
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    hint_skip_consumer_t<NPAT> cons;
    cons.tab = smem;
    cons.fin = smem + plan.fin_ofs;
    cons.acc = dfa.acc;
    cons.start = dfa.start;
    cons.pat[0] = pat0;
    cons.pat[1] = pat1;
    cons.nlines = nlines;
    cons.rc = rc;
    cons.hint = hint;
    cons.pack = pack;
    tile_pipeline_tma_early<1>(cons, &tmap, nlines, linelen,
                               smem + plan.stage_ofs + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + plan.bar_ofs) + warp * MAX_STAGES,
                               (size_t) blockIdx.x * warps_per_block + warp,
                               (size_t) gridDim.x * warps_per_block);
}

/* ---- k_dfa_lines_big -------------------------------------------------------- */

/*
 * Determinised programs whose table does not fit in shared memory (the
 * 64-pattern set: 2107 states x 48 classes x 2 bytes = 202 KB): the lines
 * still come through the TMA tile pipeline (coalesced, no L1 traffic for the
 * input), the class-compressed table is read through L1/L2 (the rows of the few
 * states a line spends its time in stay in L1).  HINT: the restart table of the
 * Pike start hint instead (entry | 0x8000 = "only the .*? thread consumed this
 * byte"), with the hint offset kept per line.
 *
 * What bounds it (ncu, profiles/r02_ncu_summary.txt): the L1/shared DATA PIPE --
 * l1tex__data_pipe_lsu_wavefronts at 94 % of its peak of one wavefront per cycle
 * and SM.  Per 32 input bytes (one byte of every lane) the pipe sees one
 * wavefront for the class look-up and ~2.4 for the table look-up (the lanes of
 * a warp sit in ~8 different states, i.e. ~11 different 32-byte sectors, of
 * which L1 serves ~4 per wavefront); the state -> state latency chain (now ONE
 * 32-bit IMAD + the load) is hidden by the 32 warps.  Measured and rejected:
 * rows padded to 128 bytes (more distinct lines: 2.13 vs 2.40 TB/s), a table of
 * 256 column addresses instead of the class bytes (3 instead of 1 wavefront per
 * class look-up: 2.0 TB/s), the start state's row in shared memory (both paths
 * execute in a mixed warp: 1.7 TB/s).  On the bench corpus 98.8 % of the
 * look-ups fall into the 64 states nearest to the start, but serving those from
 * shared memory only trades sectors for bank conflicts (~2.5 wavefronts either
 * way).
 */
template <bool HINT>
struct big_consumer_t {
    uint32_t        cls_s;      /* shared-window address of the byte-class map [256] */
    uint32_t        tab_lo, tab_hi; /* the table's address; the high word is that of every address inside
                                   it (the tables do not straddle a 4 GiB boundary: checked where they
                                   are uploaded) */
    const uint8_t  *fin;        /* global */
    uint32_t        pitch, start, acc, s, pos, p0;
    size_t          nlines;
    int32_t        *rc, *hint;
    sre_gate_pack_t pack;       /* HINT only */

    __device__ __forceinline__ void begin(size_t) { s = start; pos = 0; p0 = 0; }
    /* cls[b].  The class map stays a 256-BYTE table: text bytes then fall into distinct banks, one
     * wavefront per look-up -- the L1/shared data pipe is what saturates in this kernel
     * (l1tex__data_pipe_lsu_wavefronts 94 %); a table of 256 words costs ~3 wavefronts (measured:
     * 2.0 instead of 2.4 TB/s) */
    __device__ __forceinline__ uint32_t cls(uint32_t b) const
    {
        uint32_t c;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(cls_s + b));
        return c;
    }
    /* the state -> state chain is ONE 32-bit multiply-add (row offset + column address) and the
     * load; the column address does not depend on the state */
    __device__ __forceinline__ void step(uint32_t c, uint32_t at)
    {
        uint64_t a;
        uint16_t e16;
        uint32_t col;
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(col) : "r"(c), "r"(tab_lo));
        const uint32_t lo = s * pitch + col;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(lo), "r"(tab_hi));
        asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(e16) : "l"(a));
        const uint32_t e = e16;
        if (HINT) {
            s = e & 0x7fffu;
            p0 = (e & 0x8000u) ? at : p0;
        } else {
            s = e;
        }
    }
    __device__ __forceinline__ void byte(uint32_t b)
    {
        pos++;
        step(cls(b), pos);
    }
    struct cls4_t {
        uint32_t c0, c1, c2, c3;
    };
    __device__ __forceinline__ cls4_t classes(uint32_t w) const
    {
        cls4_t r;
        r.c0 = cls(__byte_perm(w, 0, 0x4440));
        r.c1 = cls(__byte_perm(w, 0, 0x4441));
        r.c2 = cls(__byte_perm(w, 0, 0x4442));
        r.c3 = cls(__byte_perm(w, 0, 0x4443));
        return r;
    }
    __device__ __forceinline__ void word(const cls4_t &c, uint32_t at)
    {
        step(c.c0, at + 1);
        step(c.c1, at + 2);
        step(c.c2, at + 3);
        step(c.c3, at + 4);
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        if (s != acc) {         /* absorbing: nothing can change any more (and the hint is frozen) */
            /* the class look-ups of a word are issued before the table chain of the word in front of
             * it (the asm statements are volatile: their order is the program's) */
            const cls4_t cx = classes(v.x), cy = classes(v.y);
            word(cx, pos);
            const cls4_t cz = classes(v.z);
            word(cy, pos + 4);
            const cls4_t cw = classes(v.w);
            word(cz, pos + 8);
            word(cw, pos + 12);
        }
        pos += 16;
    }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t line = group * 32 + (threadIdx.x & 31);
        bool ok = false;
        if (line < nlines) {
            ok = s == acc || __ldg(fin + s);
            rc[line] = ok ? SRE_K_OK : SRE_K_DECLINED;
            if (HINT) {
                hint[line] = (int32_t) p0;
            }
        }
        if (HINT && pack.list != nullptr) {
            gate_pack(pack, ok, line < nlines, line);
        }
    }
};

template <bool HINT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_dfa_lines_big(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, size_t nlines, uint32_t linelen,
                int32_t *__restrict__ rc, int32_t *__restrict__ hint, sre_gate_pack_t pack)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    /* [byte-class map 256][barriers][stages] */
    const uint16_t *tab = HINT ? dfa.hcls : dfa.tcls;
    {
        const uint8_t *map = HINT ? dfa.hclsmap : dfa.clsmap;
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
            smem[i] = map[i];
        }
    }
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    big_consumer_t<HINT> cons;
    cons.cls_s = smem_u32(smem);
    cons.tab_lo = (uint32_t) reinterpret_cast<uint64_t>(tab);
    cons.tab_hi = (uint32_t) (reinterpret_cast<uint64_t>(tab) >> 32);
    cons.pitch = 2 * (HINT ? dfa.hncls : dfa.nclasses);
    cons.fin = dfa.fin;
    cons.start = dfa.start;
    cons.acc = dfa.acc;
    cons.nlines = nlines;
    cons.rc = rc;
    cons.hint = hint;
    cons.pack = pack;
    tile_pipeline_tma_early<1>(cons, &tmap, nlines, linelen, smem + 4096 + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + 2048) + warp * MAX_STAGES,
                               (size_t) blockIdx.x * warps_per_block + warp, (size_t) gridDim.x * warps_per_block);
}

/* ---- k_dfa_generic --------------------------------------------------------- */

template <bool CLS, bool SMEM_TAB>
__global__ void __launch_bounds__(256)
k_dfa_generic(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
              size_t nlines, size_t pitch, size_t linelen, uint32_t *state_io, int from_init, int eof,
              int32_t *__restrict__ rc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, CLS);
    const uint8_t *tab = CLS ? reinterpret_cast<const uint8_t *>(dfa.tcls) : dfa.t256;
    const uint8_t *fin = dfa.fin, *cls = dfa.clsmap;
    /* the byte-class map always sits in shared memory: it is on every byte's
     * chain, and 256 bytes do not take anything from L1, which a table too
     * large for shared memory lives on */
    __shared__ uint8_t s_clsmap[256];
    if (CLS && !SMEM_TAB) {
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
            s_clsmap[i] = dfa.clsmap[i];
        }
        __syncthreads();
        cls = s_clsmap;
    }
    if (SMEM_TAB) {
        load_table(smem, tab, CLS ? plan.tab_bytes : (size_t) dfa.nstates * 256);
        load_table(smem + plan.fin_ofs, dfa.fin, align_up(dfa.nstates, 16));
        if (CLS) {
            load_table(smem + plan.cls_ofs, dfa.clsmap, 256);
        }
        __syncthreads();
        tab = smem;
        fin = smem + plan.fin_ofs;
        cls = smem + plan.cls_ofs;
    }
    step256_t st256 = { tab };
    stepcls_t stcls = { reinterpret_cast<const uint16_t *>(tab), cls, dfa.nclasses };

    for (size_t line = (size_t) blockIdx.x * blockDim.x + threadIdx.x; line < nlines;
         line += (size_t) gridDim.x * blockDim.x)
    {
        size_t p = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t end = offsets ? (size_t) offsets[line + 1] : p + linelen;
        uint32_t s = (from_init || state_io == nullptr) ? dfa.start : state_io[line];

        while (p < end && ((reinterpret_cast<uintptr_t>(buf) + p) & 15)) {
            const uint32_t b = buf[p++];
            s = CLS ? stcls.byte(s, b) : st256.byte(s, b);
        }
        while (p + 16 <= end) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(buf + p));
            if (CLS) {
                s = stcls.word(stcls.word(stcls.word(stcls.word(s, v.x), v.y), v.z), v.w);
            } else {
                s = st256.word(st256.word(st256.word(st256.word(s, v.x), v.y), v.z), v.w);
            }
            p += 16;
        }
        while (p < end) {
            const uint32_t b = buf[p++];
            s = CLS ? stcls.byte(s, b) : st256.byte(s, b);
        }

        int32_t r;
        if (s == dfa.acc) {
            r = SRE_K_OK;
        } else if (eof) {
            r = fin[s] ? SRE_K_OK : SRE_K_DECLINED;
        } else {
            r = SRE_K_AGAIN;
        }
        rc[line] = r;
        if (state_io) {
            state_io[line] = s;
        }
    }
}

/* ---- k_dfa_generic_hint ---------------------------------------------------- */

/* verdict + Pike start hint (see k_dfa_lines_hint) from the class-compressed
 * restart table, for DFAs of any size and lines of any alignment */
template <bool SMEM_TAB>
__global__ void __launch_bounds__(256)
k_dfa_generic_hint(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
                   size_t nlines, size_t pitch, size_t linelen, int32_t *__restrict__ rc,
                   int32_t *__restrict__ hint)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint8_t s_cls[256];
    const uint16_t *tab = dfa.hcls;
    const uint32_t C = dfa.hncls;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        s_cls[i] = dfa.hclsmap[i];
    }
    if (SMEM_TAB) {
        load_table(smem, reinterpret_cast<const uint8_t *>(dfa.hcls), align_up((size_t) dfa.nstates * C * 2, 16));
        tab = reinterpret_cast<const uint16_t *>(smem);
    }
    __syncthreads();

    for (size_t line = (size_t) blockIdx.x * blockDim.x + threadIdx.x; line < nlines;
         line += (size_t) gridDim.x * blockDim.x)
    {
        size_t p = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t begin = p, end = offsets ? (size_t) offsets[line + 1] : p + linelen;
        uint32_t s = dfa.start, p0 = 0;
        auto step = [&](uint32_t b) {
            const uint32_t e = SMEM_TAB ? tab[s * C + s_cls[b]] : __ldg(tab + s * C + s_cls[b]);
            s = e & 0x7fff;
            p++;
            if (e & 0x8000) {
                p0 = (uint32_t) (p - begin);
            }
        };
        while (p < end && ((reinterpret_cast<uintptr_t>(buf) + p) & 15) && s != dfa.acc) {
            step(buf[p]);
        }
        while (p + 16 <= end && s != dfa.acc) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(buf + p));
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    step((w[i] >> (8 * k)) & 0xff);
                }
            }
        }
        while (p < end && s != dfa.acc) {
            step(buf[p]);
        }
        rc[line] = (s == dfa.acc || dfa.fin[s]) ? SRE_K_OK : SRE_K_DECLINED;
        hint[line] = (int32_t) p0;
    }
}

/* ---- k_nfa64_lines ---------------------------------------------------------- */

/*
 * The bit-parallel NFA for programs of at most 64 lowered states whose subset
 * construction blows up (counted repetitions: /[ab]*a[ab]{15}c/ has 20 states
 * and 2^16 subsets): ONE THREAD per line, the thread set in a 64-bit register,
 * input staged by the same TMA tile pipeline as the DFA kernels.  Per byte (step
 * rule of lower/sre_lower.h):  hit |= S & mt[c];  M = S & mv[c];
 * S' = any_row | (M & shift) << 1 | OR of follow[s] over the other movers s.
 * The warp-per-line kernel below spends a whole warp on the same 64 bits.
 */
/* K = number of non-shift movers handled without a loop (0..4), -1: any number (loop over the set
 * bits: lanes diverge whenever some lane's byte moves such a state -- 20 of 32 lanes active on the
 * bench regex; with the masked ORs all 32 are) */
template <int K>
struct nfa64_consumer_t {
    const uint64_t *mv, *mt, *follow;   /* shared memory: [nclasses], [nclasses], [nkinds][64] */
    const uint8_t  *cls, *kind;         /* shared memory: [256], [nclasses]                    */
    uint64_t        init, mt_eof, shiftm, complexm, anyrow[3];
    uint64_t        S, hit;
    uint32_t        nkinds, cidx[4];
    size_t          nlines;
    int32_t        *rc;

    __device__ __forceinline__ void begin(size_t) { S = init; hit = 0; }
    __device__ __forceinline__ void byte(uint32_t b)
    {
        const uint32_t c = cls[b];
        const uint32_t k = nkinds == 1 ? 0u : kind[c];
        hit |= S & mt[c];
        const uint64_t m = S & mv[c];
        uint64_t nxt = anyrow[0];
        if (nkinds != 1) {
            nxt = k == 0 ? anyrow[0] : k == 1 ? anyrow[1] : anyrow[2];
        }
        nxt |= (m & shiftm) << 1;
        if (K >= 0) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                const uint64_t on = 0ull - ((m >> cidx[j]) & 1ull);
                nxt |= on & follow[k * 64 + cidx[j]];
            }
        } else {
            uint64_t cm = m & complexm;
            while (cm) {
                const uint32_t i = __ffsll((long long) cm) - 1;
                cm &= cm - 1;
                nxt |= follow[k * 64 + i];
            }
        }
        S = nxt;
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        if (hit != 0) {
            return;             /* a step saw a live MATCH thread: the verdict is final */
        }
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                byte((w[i] >> (8 * q)) & 0xff);
            }
        }
    }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t line = group * 32 + (threadIdx.x & 31);
        if (line < nlines) {
            rc[line] = (hit != 0 || (S & mt_eof) != 0) ? SRE_K_OK : SRE_K_DECLINED;
        }
    }
};

template <int K>
__global__ void __launch_bounds__(1024, 1)
k_nfa64_lines(sre_dev_nfa64_t nfa, const __grid_constant__ CUtensorMap tmap, size_t nlines, uint32_t linelen,
              int32_t *__restrict__ rc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    /* [mv C*8][mt C*8][follow nkinds*64*8][cls 256][kind 256][barriers][stages] */
    uint64_t *s_mv = reinterpret_cast<uint64_t *>(smem);
    uint64_t *s_mt = s_mv + 256;
    uint64_t *s_follow = s_mt + 256;
    uint8_t *s_cls = reinterpret_cast<uint8_t *>(s_follow + 3 * 64);
    uint8_t *s_kind = s_cls + 256;
    for (uint32_t i = threadIdx.x; i < nfa.nclasses; i += blockDim.x) {
        s_mv[i] = nfa.mv[i];
        s_mt[i] = nfa.mt[i];
        s_kind[i] = nfa.cls_kind[i];
    }
    for (uint32_t i = threadIdx.x; i < nfa.nkinds * 64; i += blockDim.x) {
        s_follow[i] = nfa.follow[i];
    }
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        s_cls[i] = nfa.clsmap[i];
    }
    __syncthreads();
    constexpr size_t BAR_OFS = 2 * 256 * 8 + 3 * 64 * 8 + 512, STAGE_OFS = 8192;
    static_assert(BAR_OFS + MAX_WARPS * MAX_STAGES * 8 <= STAGE_OFS, "layout");

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    nfa64_consumer_t<K> cons;
    for (int j = 0; j < 4; j++) {
        cons.cidx[j] = nfa.cidx[j];
    }
    cons.mv = s_mv;
    cons.mt = s_mt;
    cons.follow = s_follow;
    cons.cls = s_cls;
    cons.kind = s_kind;
    cons.init = nfa.init;
    cons.mt_eof = nfa.mt_eof;
    cons.shiftm = nfa.shift_mask;
    cons.complexm = nfa.complex_mask;
    cons.anyrow[0] = nfa.any_follow[0];
    cons.anyrow[1] = nfa.any_follow[1];
    cons.anyrow[2] = nfa.any_follow[2];
    cons.nkinds = nfa.nkinds;
    cons.nlines = nlines;
    cons.rc = rc;
    tile_pipeline_tma_early<1>(cons, &tmap, nlines, linelen, smem + STAGE_OFS + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + BAR_OFS) + warp * MAX_STAGES,
                               (size_t) blockIdx.x * warps_per_block + warp, (size_t) gridDim.x * warps_per_block);
}

/* ---- k_nfa_lines ----------------------------------------------------------- */

/*
 * One warp per line.  The thread set has WPL*1024 bits; lane l holds words
 * j*32 + l (j < WPL).  All bitset tables are padded to WP = 32*WPL words per row.
 * Per byte: class look-up, S & mv[class] = movers, then
 *   - the ".*?" `any` state always moves: its follow row (the fresh-start
 *     closure) is OR-ed in from registers / shared memory, no look-up;
 *   - states whose only successor is s+1 move by a one-bit shift across lanes;
 *   - the remaining ("complex") movers OR their follow rows, found by ballot.
 * Match detection: when no MATCH state carries a pending look-ahead assertion
 * the kernel only accumulates the states seen and tests them against the MATCH
 * mask every 16 bytes; otherwise it tests S & mt[class] at every step.
 * The input is read with warp-uniform 16-byte loads (one transaction per warp).
 */
template <int WPL, bool SMEM_TAB>
__global__ void __launch_bounds__(256)
k_nfa_lines(sre_dev_nfa_t nfa, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
            size_t nlines, size_t pitch, size_t linelen, uint32_t *state_io, int from_init, int eof,
            int32_t *__restrict__ rc)
{
    extern __shared__ __align__(1024) uint32_t nfa_smem[];
    __shared__ uint8_t s_cls[256];
    __shared__ uint8_t s_kind[256];
    constexpr uint32_t WP = 32 * WPL;

    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        s_cls[i] = nfa.clsmap[i];
        s_kind[i] = i < nfa.nclasses ? nfa.cls_kind[i] : 0;
    }
    const uint32_t *mvtab = nfa.mv, *mttab = nfa.mt;
    if (SMEM_TAB) {
        const uint32_t n = nfa.nclasses * WP;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            nfa_smem[i] = nfa.mv[i];
            if (nfa.match_lookahead) {
                nfa_smem[n + i] = nfa.mt[i];
            }
        }
        mvtab = nfa_smem;
        mttab = nfa_smem + n;
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const size_t warps_total = ((size_t) gridDim.x * blockDim.x) >> 5;
    const size_t gw = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool la = nfa.match_lookahead != 0;

    uint32_t shiftm[WPL], mteof[WPL], complexm[WPL], matchm[WPL];
#pragma unroll
    for (int j = 0; j < WPL; j++) {
        shiftm[j] = nfa.shift_mask[j * 32 + lane];
        mteof[j] = nfa.mt_eof[j * 32 + lane];
        complexm[j] = nfa.complex_mask[j * 32 + lane];
        matchm[j] = nfa.match_mask[j * 32 + lane];
    }

    for (size_t line = gw; line < nlines; line += warps_total) {
        size_t p = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t end = offsets ? (size_t) offsets[line + 1] : p + linelen;

        uint32_t S[WPL], seen[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            S[j] = (from_init || state_io == nullptr) ? nfa.init[j * 32 + lane]
                                                      : state_io[line * WP + j * 32 + lane];
            seen[j] = 0;
        }
        uint32_t hit = 0;

        auto step = [&](uint32_t b) {
            const uint32_t c = s_cls[b];
            const uint32_t *mvrow = mvtab + (size_t) c * WP;
            uint32_t m[WPL], nxt[WPL];
            const uint32_t *anyrow = nfa.any_follow + (nfa.nkinds == 1 ? 0 : s_kind[c]) * WP;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                if (la) {
                    hit |= S[j] & (SMEM_TAB ? mttab[(size_t) c * WP + j * 32 + lane]
                                            : __ldg(mttab + (size_t) c * WP + j * 32 + lane));
                } else {
                    seen[j] |= S[j];
                }
                m[j] = S[j] & (SMEM_TAB ? mvrow[j * 32 + lane] : __ldg(mvrow + j * 32 + lane));
                /* the `any` state is in every set and consumes every byte */
                nxt[j] = __ldg(anyrow + j * 32 + lane);
            }
            uint32_t carry = 0;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const uint32_t sh = m[j] & shiftm[j];
                uint32_t up = __shfl_up_sync(FULL, sh >> 31, 1);
                if (lane == 0) {
                    up = carry;
                }
                carry = __shfl_sync(FULL, sh >> 31, 31);
                nxt[j] |= (sh << 1) | up;
            }
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const uint32_t cm = m[j] & complexm[j];
                uint32_t bal = __ballot_sync(FULL, cm != 0);
                while (bal) {
                    const uint32_t src = __ffs(bal) - 1;
                    bal &= bal - 1;
                    uint32_t w = __shfl_sync(FULL, cm, src);
                    while (w) {
                        const uint32_t bit = __ffs(w) - 1;
                        w &= w - 1;
                        const uint32_t state = ((uint32_t) j * 32 + src) * 32 + bit;
                        const int32_t r = __ldg(nfa.rowidx + state);
                        const uint32_t kind = nfa.nkinds == 1 ? 0 : s_kind[c];
                        const uint32_t *row = nfa.follow + ((size_t) kind * nfa.nrows + r) * WP;
#pragma unroll
                        for (int jj = 0; jj < WPL; jj++) {
                            nxt[jj] |= __ldg(row + jj * 32 + lane);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                S[j] = nxt[j];
            }
        };
        auto matched_so_far = [&]() {
            uint32_t h = hit;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                h |= seen[j] & matchm[j];
            }
            return __any_sync(FULL, h) != 0;
        };

        bool done = false;
        while (p < end && ((reinterpret_cast<uintptr_t>(buf) + p) & 15)) {
            step(buf[p++]);                 /* warp-uniform address */
        }
        while (!done && p + 16 <= end) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(buf + p));
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    step((w[i] >> (8 * k)) & 0xff);
                }
            }
            p += 16;
            done = matched_so_far();
        }
        while (!done && p < end) {
            step(buf[p++]);
        }

        int32_t r;
        if (matched_so_far()) {
            r = SRE_K_OK;
        } else if (eof) {
            uint32_t h = 0;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                h |= S[j] & mteof[j];
            }
            r = __any_sync(FULL, h) ? SRE_K_OK : SRE_K_DECLINED;
        } else {
            r = SRE_K_AGAIN;
        }
        if (lane == 0) {
            rc[line] = r;
        }
        if (state_io) {
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                state_io[line * WP + j * 32 + lane] = S[j];
            }
        }
    }
}

}  // namespace

namespace sre_dev {
static int g_num_sms = 0;
/* TMA L2 promotion 256 B: measured best on B200 (5.5 -> 6.1+ TB/s on the line kernels): the
 * neighbouring 128 B of a row are the next tile of the same line */
const CUtensorMapL2promotion g_l2_promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;

cudaError_t make_row_tensor_map(CUtensorMap *map, const uint8_t *buf, size_t nrows, size_t pitch, int tw)
{
    typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
    static encode_fn_t encode = nullptr;
    if (encode == nullptr) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
            return e != cudaSuccess ? e : cudaErrorNotSupported;
        }
        encode = reinterpret_cast<encode_fn_t>(fn);
    }
    if (pitch % 16 || (reinterpret_cast<uintptr_t>(buf) & 15) || nrows == 0 || nrows >= (1ull << 31)
        || pitch >= (1ull << 31))
    {
        return cudaErrorInvalidValue;
    }
    const cuuint64_t gdim[2] = { (cuuint64_t) pitch, (cuuint64_t) nrows };
    const cuuint64_t gstride[1] = { (cuuint64_t) pitch };
    const cuuint32_t box[2] = { (cuuint32_t) tw, 32 };
    const cuuint32_t estride[2] = { 1, 1 };
    const CUtensorMapSwizzle swz = tw >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : tw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : tw == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(buf), gdim, gstride,
                              box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                              g_l2_promotion, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

int num_sms()
{
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) {
            g_num_sms = 148;
        }
    }
    return g_num_sms;
}
}  // namespace sre_dev

namespace {

template <int STAGES, bool CLS, int THREADS, int BLOCKS>
cudaError_t launch_dfa_lines_early_t(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *rc, cudaStream_t stream)
{
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, CLS);
    const int warps = THREADS / 32;
    const size_t smem = plan.stage_ofs + (size_t) warps * STAGES * 32 * 128;
    if (BLOCKS * (smem + 1024) > SMEM_LIMIT) {
        return cudaErrorInvalidConfiguration;
    }
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, buf, nlines, pitch, 128);
    if (err != cudaSuccess) {
        return err;
    }
    auto kern = k_dfa_lines_tma_early<STAGES, CLS, THREADS, BLOCKS>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (err != cudaSuccess) {
            return err;
        }
        smem_set = smem;
    }
    const size_t ngroups = (nlines + 31) / 32;
    size_t grid = (size_t) num_sms() * BLOCKS;
    const size_t need = (ngroups + warps - 1) / warps;
    if (grid > need) {
        grid = need;
    }
    kern<<<(unsigned) grid, THREADS, smem, stream>>>(dfa, tmap, nlines, (uint32_t) linelen, rc);
    return cudaGetLastError();
}


/* ---- k_nfa_packed ----------------------------------------------------------- */

/*
 * The same step for programs with at most four non-shift movers (the rule for
 * regexes whose bounded repetitions defeat determinisation: one loop state and a
 * run of shift states), with everything a byte needs in ONE shared-memory entry
 * per byte class: { mv, mt, the ".*?" row of the class's kind, the follow rows of
 * the K movers for that kind }.  Per byte: the class (LDS.U8), the entry (one or
 * two LDS.128, broadcast: text has few classes), and straight-line logic -- no
 * kind look-up, no loop over set bits, no divergence.  W = uint32_t when the
 * program has at most 32 lowered states (every 64-bit operation of k_nfa64_lines
 * is two instructions), else uint64_t.
 */
template <typename W, int K>
struct nfa_packed_consumer_t {
    static constexpr int NW = 3 + K;                                   /* words per entry */
    static constexpr int NV = (NW * (int) sizeof(W) + 15) / 16;         /* uint4 per entry */
    uint32_t        ent_s, cls_s;       /* shared-window addresses: entries [nclasses], class map [256] */
    W               init, mt_eof, shiftm, cbit[K > 0 ? K : 1];
    W               S, hit;
    size_t          nlines;
    int32_t        *rc;

    __device__ __forceinline__ void begin(size_t) { S = init; hit = 0; }
    __device__ __forceinline__ void byte(uint32_t b)
    {
        uint32_t c;
        asm("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(cls_s + b));
        union {
            uint4 v[NV];
            W     w[NV * 16 / sizeof(W)];
        } e;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                : "=r"(e.v[i].x), "=r"(e.v[i].y), "=r"(e.v[i].z), "=r"(e.v[i].w)
                : "r"(ent_s + c * (uint32_t) (NV * 16) + (uint32_t) i * 16));
        }
        hit |= S & e.w[1];
        const W m = S & e.w[0];
        W nxt = e.w[2] | (W) ((m & shiftm) << 1);
#pragma unroll
        for (int j = 0; j < K; j++) {
            nxt |= (m & cbit[j]) ? e.w[3 + j] : (W) 0;
        }
        S = nxt;
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        if (hit != 0) {
            return;             /* a step saw a live MATCH thread: the verdict is final */
        }
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                byte(__byte_perm(w[i], 0, 0x4440 + q));
            }
        }
    }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t line = group * 32 + (threadIdx.x & 31);
        if (line < nlines) {
            rc[line] = (hit != 0 || (S & mt_eof) != 0) ? SRE_K_OK : SRE_K_DECLINED;
        }
    }
};

template <typename W, int K>
__global__ void __launch_bounds__(1024, 1)
k_nfa_packed(sre_dev_nfa64_t nfa, const __grid_constant__ CUtensorMap tmap, size_t nlines, uint32_t linelen,
             int32_t *__restrict__ rc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    using cons_t = nfa_packed_consumer_t<W, K>;
    /* [entries 256 x NV x 16][cls 256][barriers][stages] */
    constexpr size_t ENT_BYTES = (size_t) 256 * cons_t::NV * 16, CLS_OFS = ENT_BYTES, BAR_OFS = CLS_OFS + 256,
                     STAGE_OFS = (BAR_OFS + MAX_WARPS * MAX_STAGES * 8 + 1023) / 1024 * 1024;
    W *ent = reinterpret_cast<W *>(smem);
    constexpr int EW = cons_t::NV * 16 / (int) sizeof(W);
    for (uint32_t c = threadIdx.x; c < nfa.nclasses; c += blockDim.x) {
        const uint32_t k = nfa.nkinds == 1 ? 0u : nfa.cls_kind[c];
        W *e = ent + (size_t) c * EW;
        e[0] = (W) nfa.mv[c];
        e[1] = (W) nfa.mt[c];
        e[2] = (W) nfa.any_follow[k];
#pragma unroll
        for (int j = 0; j < K; j++) {
            e[3 + j] = (W) nfa.follow[k * 64 + nfa.cidx[j]];
        }
    }
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        smem[CLS_OFS + i] = nfa.clsmap[i];
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    cons_t cons;
    cons.ent_s = smem_u32(smem);
    cons.cls_s = smem_u32(smem + CLS_OFS);
    cons.init = (W) nfa.init;
    cons.mt_eof = (W) nfa.mt_eof;
    cons.shiftm = (W) nfa.shift_mask;
#pragma unroll
    for (int j = 0; j < K; j++) {
        cons.cbit[j] = (W) ((uint64_t) 1 << nfa.cidx[j]);
    }
    cons.nlines = nlines;
    cons.rc = rc;
    tile_pipeline_tma_early<1>(cons, &tmap, nlines, linelen, smem + STAGE_OFS + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + BAR_OFS) + warp * MAX_STAGES,
                               (size_t) blockIdx.x * warps_per_block + warp, (size_t) gridDim.x * warps_per_block);
}

template <typename W, int K>
static cudaError_t launch_nfa_packed_t(const sre_dev_nfa64_t &nfa, const CUtensorMap &tmap, size_t nlines,
    size_t linelen, int32_t *rc, size_t grid, cudaStream_t stream)
{
    constexpr size_t NV = nfa_packed_consumer_t<W, K>::NV;
    const size_t smem = (256 * NV * 16 + 256 + MAX_WARPS * MAX_STAGES * 8 + 1023) / 1024 * 1024 + (size_t) 32 * 32 * 128;
    auto kern = k_nfa_packed<W, K>;
    static bool attr_set = false;
    if (!attr_set) {
        const cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (err != cudaSuccess) {
            return err;
        }
        attr_set = true;
    }
    kern<<<(unsigned) grid, 1024, smem, stream>>>(nfa, tmap, nlines, (uint32_t) linelen, rc);
    return cudaGetLastError();
}

}  // namespace

size_t sre_dfa_smem_table_bytes(uint32_t nstates, uint32_t nclasses, bool cls)
{
    return dfa_smem_plan(nstates, nclasses, cls).stage_ofs;
}

cudaError_t sre_launch_dfa_lines(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *rc, int variant, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    const bool cls = dfa.t256 == nullptr;
    if (launches) {
        ++*launches;
    }
    /* <stages, threads/block, blocks/SM> */
#define SRE_EARLY(ST, TH, BL)                                                                  \
    (cls ? launch_dfa_lines_early_t<ST, true, TH, BL>(dfa, buf, nlines, pitch, linelen, rc, stream) \
         : launch_dfa_lines_early_t<ST, false, TH, BL>(dfa, buf, nlines, pitch, linelen, rc, stream))
    switch (variant) {
    case 0:  return SRE_EARLY(1, 1024, 1);      /* 32 warps/SM: best measured on B200 */
    case 1:  return SRE_EARLY(1, 768, 2);       /* 48 warps/SM */
    case 2:  return SRE_EARLY(2, 768, 1);       /* 24 warps/SM, 2 stages */
    default: return cudaErrorInvalidValue;
    }
#undef SRE_EARLY
}

template <int NPAT, int THREADS, int BLOCKS>
static cudaError_t launch_dfa_lines_skip_t(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *rc, const uint32_t *pats, cudaStream_t stream)
{
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, false);
    const int warps = THREADS / 32;
    const size_t smem = plan.stage_ofs + (size_t) warps * 32 * 128;
    if (BLOCKS * (smem + 1024) > SMEM_LIMIT) {
        return cudaErrorInvalidConfiguration;
    }
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, buf, nlines, pitch, 128);
    if (err != cudaSuccess) {
        return err;
    }
    auto kern = k_dfa_lines_skipw<NPAT, THREADS, BLOCKS>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (err != cudaSuccess) {
            return err;
        }
        smem_set = smem;
    }
    const size_t ngroups = (nlines + 31) / 32;
    size_t grid = (size_t) num_sms() * BLOCKS;
    const size_t need = (ngroups + warps - 1) / warps;
    if (grid > need) {
        grid = need;
    }
    kern<<<(unsigned) grid, THREADS, smem, stream>>>(dfa, tmap, nlines, (uint32_t) linelen,
                                                   make_uint4(pats[0], pats[1], pats[2], pats[3]), rc);
    return cudaGetLastError();
}

/* npat leave bytes (1..4), replicated into all four byte lanes of pats[] */
cudaError_t sre_launch_dfa_lines_skip(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *rc, const uint32_t *pats, int npat, int variant,
    cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (dfa.t256 == nullptr || npat < 1 || npat > 4) {
        return cudaErrorInvalidValue;
    }
    if (launches) {
        ++*launches;
    }
#define SRE_SKIP(TH, BL)                                                                              \
    (npat == 1 ? launch_dfa_lines_skip_t<1, TH, BL>(dfa, buf, nlines, pitch, linelen, rc, pats, stream) \
     : npat == 2 ? launch_dfa_lines_skip_t<2, TH, BL>(dfa, buf, nlines, pitch, linelen, rc, pats, stream) \
                 : launch_dfa_lines_skip_t<4, TH, BL>(dfa, buf, nlines, pitch, linelen, rc, pats, stream))
    switch (variant) {
    case 0:  return SRE_SKIP(1024, 1);      /* best measured on B200: 94.9 % of the HBM copy peak */
    case 1:  return SRE_SKIP(768, 1);
    case 2:  return SRE_SKIP(640, 2);
    default: return cudaErrorInvalidValue;
    }
#undef SRE_SKIP
}

cudaError_t sre_launch_dfa_generic_hint(const sre_dev_dfa_t &dfa, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, int32_t *rc, int32_t *hint, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (dfa.hcls == nullptr) {
        return cudaErrorInvalidValue;
    }
    const size_t tab = align_up((size_t) dfa.nstates * dfa.hncls * 2, 16);
    const bool fits = tab <= 160 * 1024;
    size_t grid = (nlines + 255) / 256;
    const size_t cap = (size_t) num_sms() * (fits ? 1 : 8);
    if (grid > cap) {
        grid = cap;
    }
    if (launches) {
        ++*launches;
    }
    if (fits) {
        static size_t smem_set = 0;
        if (tab > smem_set) {
            cudaError_t e = cudaFuncSetAttribute(k_dfa_generic_hint<true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tab);
            if (e != cudaSuccess) {
                return e;
            }
            smem_set = tab;
        }
        k_dfa_generic_hint<true><<<(unsigned) grid, 256, tab, stream>>>(dfa, buf, offsets, nlines, pitch,
                                                                        linelen, rc, hint);
    } else {
        k_dfa_generic_hint<false><<<(unsigned) grid, 256, 0, stream>>>(dfa, buf, offsets, nlines, pitch,
                                                                       linelen, rc, hint);
    }
    return cudaGetLastError();
}

cudaError_t sre_launch_dfa_lines_hint(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *rc, int32_t *hint, const uint32_t *pats, int npat,
    sre_gate_pack_t pack, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (dfa.h256 == nullptr) {
        return cudaErrorInvalidValue;
    }
    const dfa_smem_plan_t plan = dfa_smem_plan(256, 0, false);
    const int warps = 32;
    const size_t smem = plan.stage_ofs + (size_t) warps * 32 * 128;
    const size_t smem_plain = HINT_STAGE_OFS + (size_t) warps * 32 * 128;      /* padded rows */
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, buf, nlines, pitch, 128);
    if (err != cudaSuccess) {
        return err;
    }
    static bool attr_set = false;
    if (!attr_set) {
        err = cudaFuncSetAttribute(k_dfa_lines_hint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_plain);
        if (err == cudaSuccess) {
            err = cudaFuncSetAttribute(k_dfa_lines_hint_skip<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) smem);
        }
        if (err == cudaSuccess) {
            err = cudaFuncSetAttribute(k_dfa_lines_hint_skip<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) smem);
        }
        if (err != cudaSuccess) {
            return err;
        }
        attr_set = true;
    }
    const size_t ngroups = (nlines + 31) / 32;
    size_t grid = (size_t) num_sms();
    const size_t need = (ngroups + warps - 1) / warps;
    if (grid > need) {
        grid = need;
    }
    if (launches) {
        ++*launches;
    }
    /* word skip when the start state is left by one or two byte values */
    if (pats != nullptr && npat == 1) {
        k_dfa_lines_hint_skip<1><<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, tmap, nlines, (uint32_t) linelen,
                                                                               pats[0], pats[0], rc, hint, pack);
    } else if (pats != nullptr && npat == 2) {
        k_dfa_lines_hint_skip<2><<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, tmap, nlines, (uint32_t) linelen,
                                                                               pats[0], pats[1], rc, hint, pack);
    } else {
        k_dfa_lines_hint<<<(unsigned) grid, warps * 32, smem_plain, stream>>>(dfa, tmap, nlines, (uint32_t) linelen,
                                                                            rc, hint, pack);
    }
    return cudaGetLastError();
}

static cudaError_t launch_dfa_generic(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, uint32_t *state_io,
    int from_init, int eof, int32_t *rc, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    const bool cls = dfa.t256 == nullptr;
    const size_t smem = dfa_smem_plan(dfa.nstates, dfa.nclasses, cls).stage_ofs;
    const bool fits = smem <= 96 * 1024;
    size_t grid = (nlines + 255) / 256;
    const size_t cap = (size_t) num_sms() * 8;
    if (grid > cap) {
        grid = cap;
    }
    if (launches) {
        ++*launches;
    }
#define SRE_GENERIC(CLS, SM)                                                                     \
    do {                                                                                         \
        auto kern = k_dfa_generic<CLS, SM>;                                                      \
        if (SM) {                                                                                \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int) smem);                                    \
            if (e != cudaSuccess) return e;                                                      \
        }                                                                                        \
        kern<<<(unsigned) grid, 256, SM ? smem : 0, stream>>>(dfa, buf, offsets, nlines, pitch,  \
                                                              linelen, state_io, from_init, eof, rc); \
    } while (0)
    if (cls) {
        if (fits) SRE_GENERIC(true, true); else SRE_GENERIC(true, false);
    } else {
        if (fits) SRE_GENERIC(false, true); else SRE_GENERIC(false, false);
    }
#undef SRE_GENERIC
    return cudaGetLastError();
}

cudaError_t sre_launch_dfa_ragged(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, int32_t *rc,
    cudaStream_t stream, int *launches)
{
    return launch_dfa_generic(dfa, buf, offsets, nlines, pitch, linelen, nullptr, 1, 1, rc, stream,
                              launches);
}

cudaError_t sre_launch_dfa_carry(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, uint32_t *state_io,
    int from_init, int eof, int32_t *rc, cudaStream_t stream, int *launches)
{
    return launch_dfa_generic(dfa, buf, offsets, nlines, pitch, linelen, state_io, from_init, eof, rc,
                              stream, launches);
}

cudaError_t sre_launch_nfa_lines(const sre_dev_nfa_t &nfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, uint32_t *state_io,
    int from_init, int eof, int32_t *rc, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    const uint32_t wpl = (nfa.nwords + 31) / 32;
    size_t grid = (nlines + 7) / 8;     /* 8 warps per block */
    const size_t cap = (size_t) num_sms() * 8;
    if (grid > cap) {
        grid = cap;
    }
    if (launches) {
        ++*launches;
    }
    /* mv (and mt when needed) tables in shared memory when they fit */
    const size_t tab = (size_t) nfa.nclasses * nfa.nwords * 4 * (nfa.match_lookahead ? 2 : 1);
    const bool fits = tab <= 24 * 1024;
#define SRE_NFA(W)                                                                                   \
    do {                                                                                             \
        if (fits) {                                                                                  \
            k_nfa_lines<W, true><<<(unsigned) grid, 256, tab, stream>>>(nfa, buf, offsets, nlines, pitch, \
                                                                        linelen, state_io, from_init, eof, rc); \
        } else {                                                                                     \
            k_nfa_lines<W, false><<<(unsigned) grid, 256, 0, stream>>>(nfa, buf, offsets, nlines, pitch,  \
                                                                       linelen, state_io, from_init, eof, rc); \
        }                                                                                            \
    } while (0)
    if (wpl <= 1) {
        SRE_NFA(1);
    } else if (wpl <= 2) {
        SRE_NFA(2);
    } else if (wpl <= 4) {
        SRE_NFA(4);
    } else {
        return cudaErrorInvalidValue;
    }
#undef SRE_NFA
    return cudaGetLastError();
}

/* thread-per-line NFA for <= 64 lowered states; 16-byte aligned fixed-pitch lines */
cudaError_t sre_launch_nfa64_lines(const sre_dev_nfa64_t &nfa, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int32_t *rc, cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    const int warps = 32;
    const size_t smem = 8192 + (size_t) warps * 32 * 128;
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, buf, nlines, pitch, 128);
    if (err != cudaSuccess) {
        return err;
    }
    if (nfa.ncomplex <= 4) {
        const size_t ngroups = (nlines + 31) / 32;
        size_t grid = (size_t) num_sms();
        const size_t need = (ngroups + warps - 1) / warps;
        if (grid > need) {
            grid = need;
        }
        if (launches) {
            ++*launches;
        }
        const bool w32 = nfa.nstates <= 32;
        switch (nfa.ncomplex) {
#define SRE_PACKED(KK)                                                                                       \
        case KK: return w32 ? launch_nfa_packed_t<uint32_t, KK>(nfa, tmap, nlines, linelen, rc, grid, stream) \
                            : launch_nfa_packed_t<uint64_t, KK>(nfa, tmap, nlines, linelen, rc, grid, stream);
        SRE_PACKED(0)
        SRE_PACKED(1)
        SRE_PACKED(2)
        SRE_PACKED(3)
        SRE_PACKED(4)
#undef SRE_PACKED
        }
    }
    typedef void (*kern_t)(sre_dev_nfa64_t, const CUtensorMap, size_t, uint32_t, int32_t *);
    static const kern_t kerns[6] = { k_nfa64_lines<-1>, k_nfa64_lines<0>, k_nfa64_lines<1>, k_nfa64_lines<2>,
                                     k_nfa64_lines<3>, k_nfa64_lines<4> };
    const int ki = nfa.ncomplex <= 4 ? (int) nfa.ncomplex + 1 : 0;
    static bool attr_set[6];
    if (!attr_set[ki]) {
        err = cudaFuncSetAttribute(kerns[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (err != cudaSuccess) {
            return err;
        }
        attr_set[ki] = true;
    }
    const size_t ngroups = (nlines + 31) / 32;
    size_t grid = (size_t) num_sms();
    const size_t need = (ngroups + warps - 1) / warps;
    if (grid > need) {
        grid = need;
    }
    if (launches) {
        ++*launches;
    }
    kerns[ki]<<<(unsigned) grid, warps * 32, smem, stream>>>(nfa, tmap, nlines, (uint32_t) linelen, rc);
    return cudaGetLastError();
}

/* tiled input, class table through L1/L2: verdict (and, with hint != NULL, the Pike start hint)
 * for DFAs whose table exceeds shared memory; 16-byte aligned fixed-pitch lines */
template <bool HINT, int WARPS>
static cudaError_t launch_dfa_lines_big_t(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int32_t *rc, int32_t *hint, sre_gate_pack_t pack, cudaStream_t stream)
{
    const size_t smem = 4096 + (size_t) WARPS * 32 * 128;
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, buf, nlines, pitch, 128);
    if (err != cudaSuccess) {
        return err;
    }
    auto kern = k_dfa_lines_big<HINT, WARPS>;
    static bool attr_set = false;
    if (!attr_set) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (err != cudaSuccess) {
            return err;
        }
        attr_set = true;
    }
    const size_t ngroups = (nlines + 31) / 32;
    size_t grid = (size_t) num_sms();
    const size_t need = (ngroups + WARPS - 1) / WARPS;
    if (grid > need) {
        grid = need;
    }
    kern<<<(unsigned) grid, WARPS * 32, smem, stream>>>(dfa, tmap, nlines, (uint32_t) linelen, rc, hint, pack);
    return cudaGetLastError();
}

cudaError_t sre_launch_dfa_lines_big(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int32_t *rc, int32_t *hint, sre_gate_pack_t pack, int variant, cudaStream_t stream,
    int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (hint != nullptr && dfa.hcls == nullptr) {
        return cudaErrorInvalidValue;
    }
    if (launches) {
        ++*launches;
    }
    /* 32 warps per SM (128 KB of staging, ~100 KB left to L1: measured best) or 16 (64 KB, more L1) */
    if (hint != nullptr) {
        return variant == 1 ? launch_dfa_lines_big_t<true, 16>(dfa, buf, nlines, pitch, linelen, rc, hint, pack, stream)
                            : launch_dfa_lines_big_t<true, 32>(dfa, buf, nlines, pitch, linelen, rc, hint, pack, stream);
    }
    return variant == 1 ? launch_dfa_lines_big_t<false, 16>(dfa, buf, nlines, pitch, linelen, rc, hint, pack, stream)
                        : launch_dfa_lines_big_t<false, 32>(dfa, buf, nlines, pitch, linelen, rc, hint, pack, stream);
}
