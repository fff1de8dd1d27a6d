/*
 * sre_thompson.cu -- Thompson (boolean) match kernels for sm_100a.
 *
 * What they replace: sre_vm_thompson_exec (reference sre_vm_thompson.c:63-270)
 * and the code its x86-64 JIT emits (sre_vm_thompson_x64.dasc:746-886).  The
 * thread-list loop becomes table look-ups over the lowered program
 * (../lower/sre_lower.h):
 *
 *   k_dfa_lines    one CUDA thread per line, determinised program.  Input
 *                  tiles (32 lines x TW bytes per warp) are staged through
 *                  shared memory with 16-byte cp.async copies (coalesced: the
 *                  lanes that share a row cover a contiguous 64/128-byte
 *                  segment), XOR-swizzled so that the per-lane 16-byte reads of
 *                  "my row" are bank-conflict free.  Per input byte the inner
 *                  loop is one PRMT (splice the byte under the state) and one
 *                  LDS.U8 into the [state][byte] table.  HBM-bound by design.
 *   k_dfa_generic  same automaton, any alignment / ragged offsets, optional
 *                  state carry (SRE_AGAIN) -- correctness tier.
 *   k_nfa_lines    the general tier: one warp per line, the NFA thread set is a
 *                  warp-wide bitmask (lane l owns words l, l+32, ... of it, up
 *                  to 4096 states), successor sets come from a shift for
 *                  "next pc" states and from follow-row ORs for the rest.
 */
#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

/* ---- k_dfa_lines ----------------------------------------------------------- */

template <bool CLS>
struct line_consumer_t {
    step256_t       st256;
    stepcls_t       stcls;
    const uint8_t  *fin;
    uint32_t        start, acc, s;
    size_t          nlines;
    int32_t        *rc;

    __device__ __forceinline__ void begin() { s = start; }
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

template <int TW, int STAGES, bool CLS>
__global__ void __launch_bounds__(1024, 1)
k_dfa_lines(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t nlines, size_t pitch,
            uint32_t linelen, int32_t *__restrict__ rc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, CLS);
    uint8_t *s_tab = smem;
    uint8_t *s_fin = smem + plan.fin_ofs;
    uint8_t *s_cls = smem + plan.cls_ofs;

    load_table(s_tab, CLS ? reinterpret_cast<const uint8_t *>(dfa.tcls) : dfa.t256, plan.tab_bytes);
    load_table(s_fin, dfa.fin, align_up(dfa.nstates, 16));
    if (CLS) {
        load_table(s_cls, dfa.clsmap, 256);
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    line_consumer_t<CLS> cons;
    cons.st256.tab = s_tab;
    cons.stcls.tab = reinterpret_cast<const uint16_t *>(s_tab);
    cons.stcls.cls = s_cls;
    cons.stcls.ncls = dfa.nclasses;
    cons.fin = s_fin;
    cons.start = dfa.start;
    cons.acc = dfa.acc;
    cons.nlines = nlines;
    cons.rc = rc;

    tile_pipeline<TW, STAGES>(cons, buf, nlines, pitch, linelen,
                              smem + plan.stage_ofs + (size_t) warp * STAGES * 32 * TW,
                              (size_t) blockIdx.x * warps_per_block + warp,
                              (size_t) gridDim.x * warps_per_block);
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
    if (SMEM_TAB) {
        load_table(smem, tab, plan.tab_bytes);
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

/* ---- k_nfa_lines ----------------------------------------------------------- */

/*
 * One warp per line.  The thread set has WPL*1024 bits; lane l holds words
 * j*32 + l (j < WPL).  All bitset tables are padded to 32*WPL words per row.
 */
template <int WPL>
__global__ void __launch_bounds__(256)
k_nfa_lines(sre_dev_nfa_t nfa, const uint8_t *__restrict__ buf, const int64_t *__restrict__ offsets,
            size_t nlines, size_t pitch, size_t linelen, uint32_t *state_io, int from_init, int eof,
            int32_t *__restrict__ rc)
{
    __shared__ uint8_t s_cls[256];
    __shared__ uint8_t s_kind[256];
    constexpr uint32_t WP = 32 * WPL;

    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        s_cls[i] = nfa.clsmap[i];
        s_kind[i] = i < nfa.nclasses ? nfa.cls_kind[i] : 0;
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const size_t warps_total = ((size_t) gridDim.x * blockDim.x) >> 5;
    const size_t gw = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;

    uint32_t shiftm[WPL], mteof[WPL];
#pragma unroll
    for (int j = 0; j < WPL; j++) {
        shiftm[j] = nfa.shift_mask[j * 32 + lane];
        mteof[j] = nfa.mt_eof[j * 32 + lane];
    }

    for (size_t line = gw; line < nlines; line += warps_total) {
        size_t p = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t end = offsets ? (size_t) offsets[line + 1] : p + linelen;

        uint32_t S[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            S[j] = (from_init || state_io == nullptr) ? nfa.init[j * 32 + lane]
                                                      : state_io[line * WP + j * 32 + lane];
        }
        uint32_t hit = 0;

        while (p < end && !__any_sync(FULL, hit)) {
            const uint32_t n = (uint32_t) (end - p < 32 ? end - p : 32);
            const uint32_t mine = lane < n ? buf[p + lane] : 0;

            for (uint32_t i = 0; i < n; i++) {
                const uint32_t b = __shfl_sync(FULL, mine, i);
                const uint32_t c = s_cls[b], kind = s_kind[c];
                const uint32_t *mvrow = nfa.mv + (size_t) c * WP, *mtrow = nfa.mt + (size_t) c * WP;
                uint32_t m[WPL], nxt[WPL];
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    hit |= S[j] & __ldg(mtrow + j * 32 + lane);
                    m[j] = S[j] & __ldg(mvrow + j * 32 + lane);
                }
                /* states whose only successor is s+1: shift left by one bit */
                uint32_t carry = 0;
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    const uint32_t sh = m[j] & shiftm[j];
                    uint32_t up = __shfl_up_sync(FULL, sh >> 31, 1);
                    if (lane == 0) {
                        up = carry;
                    }
                    carry = __shfl_sync(FULL, sh >> 31, 31);
                    nxt[j] = (sh << 1) | up;
                }
                /* the rest: OR their follow rows */
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    const uint32_t cm = m[j] & ~shiftm[j];
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
            }
            p += n;
        }

        int32_t r;
        if (__any_sync(FULL, hit)) {
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

template <int TW, int STAGES, bool CLS>
cudaError_t launch_dfa_lines_t(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *rc, cudaStream_t stream)
{
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, CLS);
    const size_t per_warp = (size_t) STAGES * 32 * TW;
    /* prefer 2 blocks of 16 warps per SM; fall back to what fits */
    int warps = 16, blocks_per_sm = 2;
    while (blocks_per_sm * (plan.stage_ofs + warps * per_warp + 1024) > SMEM_LIMIT) {
        if (blocks_per_sm == 2) {
            blocks_per_sm = 1;
            warps = 32;
        } else if (warps > 4) {
            warps -= 4;
        } else {
            return cudaErrorInvalidConfiguration;
        }
    }
    const size_t smem = plan.stage_ofs + warps * per_warp;
    auto kern = k_dfa_lines<TW, STAGES, CLS>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (err != cudaSuccess) {
        return err;
    }
    const size_t ngroups = (nlines + 31) / 32;
    size_t grid = (size_t) num_sms() * blocks_per_sm;
    const size_t need = (ngroups + warps - 1) / warps;
    if (grid > need) {
        grid = need;
    }
    kern<<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, buf, nlines, pitch, (uint32_t) linelen, rc);
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
    switch (variant) {
    case 1:
        return cls ? launch_dfa_lines_t<128, 2, true>(dfa, buf, nlines, pitch, linelen, rc, stream)
                   : launch_dfa_lines_t<128, 2, false>(dfa, buf, nlines, pitch, linelen, rc, stream);
    case 2:
        return cls ? launch_dfa_lines_t<32, 4, true>(dfa, buf, nlines, pitch, linelen, rc, stream)
                   : launch_dfa_lines_t<32, 4, false>(dfa, buf, nlines, pitch, linelen, rc, stream);
    case 3:
        return cls ? launch_dfa_lines_t<64, 4, true>(dfa, buf, nlines, pitch, linelen, rc, stream)
                   : launch_dfa_lines_t<64, 4, false>(dfa, buf, nlines, pitch, linelen, rc, stream);
    default:
        return cls ? launch_dfa_lines_t<64, 3, true>(dfa, buf, nlines, pitch, linelen, rc, stream)
                   : launch_dfa_lines_t<64, 3, false>(dfa, buf, nlines, pitch, linelen, rc, stream);
    }
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
    if (wpl <= 1) {
        k_nfa_lines<1><<<(unsigned) grid, 256, 0, stream>>>(nfa, buf, offsets, nlines, pitch, linelen,
                                                           state_io, from_init, eof, rc);
    } else if (wpl <= 2) {
        k_nfa_lines<2><<<(unsigned) grid, 256, 0, stream>>>(nfa, buf, offsets, nlines, pitch, linelen,
                                                           state_io, from_init, eof, rc);
    } else if (wpl <= 4) {
        k_nfa_lines<4><<<(unsigned) grid, 256, 0, stream>>>(nfa, buf, offsets, nlines, pitch, linelen,
                                                           state_io, from_init, eof, rc);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
