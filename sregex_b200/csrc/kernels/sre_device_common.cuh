/*
 * sre_device_common.cuh -- device helpers shared by the Thompson kernels: DFA
 * step functors, the shared-memory plan of the DFA tables, TMA / mbarrier
 * wrappers and the per-warp tile pipeline that stages 32 rows x 128 bytes
 * through shared memory (rows = lines for the k_dfa_lines family, stream pieces
 * for k_stream_pieces).
 */
#ifndef SRE_DEVICE_COMMON_CUH
#define SRE_DEVICE_COMMON_CUH

#include <cuda.h>
#include "sre_kernels.cuh"

namespace sre_dev {

constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ constexpr size_t align_up(size_t v, size_t a)
{
    return (v + a - 1) / a * a;
}

/* ---- DFA step functors ---------------------------------------------------- */

/* [state][byte] u8 table: one PRMT + one LDS.U8 per input byte */
struct step256_t {
    const uint8_t *tab;
    __device__ __forceinline__ uint32_t word(uint32_t s, uint32_t w) const
    {
        s = tab[__byte_perm(w, s, 0x5540)];
        s = tab[__byte_perm(w, s, 0x5541)];
        s = tab[__byte_perm(w, s, 0x5542)];
        s = tab[__byte_perm(w, s, 0x5543)];
        return s;
    }
    __device__ __forceinline__ uint32_t byte(uint32_t s, uint32_t b) const
    {
        return tab[(s << 8) | b];
    }
};

/*
 * The same table with rows padded to 260 bytes (65 words): row s starts s
 * banks further on, so lanes that sit in different states and read bytes of the
 * same 4-byte group -- the rule on text over a small alphabet -- no longer
 * collide on one bank.  Costs one IMAD per byte on top of the byte extract.
 */
constexpr uint32_t ROW260 = 260;
struct step260_t {
    uint32_t tab_s;             /* shared-window address of the table; its low byte must be 0 */
    __device__ __forceinline__ static uint32_t lds_u8(uint32_t addr)
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
        return v;
    }
    /* The byte's address within row 0 is ONE PRMT (the byte spliced under the upper three bytes of
     * the table's address), off the state -> state chain, which is then one IMAD and one LDS per
     * byte (the IMAD on the fma pipe, which these kernels leave idle). */
    __device__ __forceinline__ uint32_t word(uint32_t s, uint32_t w) const
    {
        const uint32_t a0 = __byte_perm(w, tab_s, 0x7650), a1 = __byte_perm(w, tab_s, 0x7651);
        const uint32_t a2 = __byte_perm(w, tab_s, 0x7652), a3 = __byte_perm(w, tab_s, 0x7653);
        s = lds_u8(s * ROW260 + a0);
        s = lds_u8(s * ROW260 + a1);
        s = lds_u8(s * ROW260 + a2);
        s = lds_u8(s * ROW260 + a3);
        return s;
    }
    __device__ __forceinline__ uint32_t byte(uint32_t s, uint32_t b) const
    {
        return lds_u8(s * ROW260 + tab_s + b);
    }
};

/* [nstates][256] in global memory -> rows of ROW260 bytes in shared memory */
__device__ __forceinline__ void load_table260(uint8_t *dst, const uint8_t *src, uint32_t nstates)
{
    for (uint32_t i = threadIdx.x; i < nstates * 64; i += blockDim.x) {
        reinterpret_cast<uint32_t *>(dst)[(i >> 6) * (ROW260 / 4) + (i & 63)] =
            reinterpret_cast<const uint32_t *>(src)[i];
    }
}

/* byte-class compressed u16 table: extract, LDS.U8 class, IMAD, LDS.U16 */
struct stepcls_t {
    const uint16_t *tab;
    const uint8_t  *cls;
    uint32_t        ncls;
    __device__ __forceinline__ uint32_t byte(uint32_t s, uint32_t b) const
    {
        return tab[s * ncls + cls[b]];
    }
    __device__ __forceinline__ uint32_t word(uint32_t s, uint32_t w) const
    {
        s = byte(s, w & 0xff);
        s = byte(s, (w >> 8) & 0xff);
        s = byte(s, (w >> 16) & 0xff);
        s = byte(s, w >> 24);
        return s;
    }
};

/* cooperative copy of a device table into shared memory (16-byte granules;
 * device tables are allocated padded to 16 bytes) */
__device__ __forceinline__ void load_table(uint8_t *dst, const uint8_t *src, size_t bytes)
{
    for (size_t i = (size_t) threadIdx.x * 16; i < bytes; i += (size_t) blockDim.x * 16) {
        *reinterpret_cast<uint4 *>(dst + i) = *reinterpret_cast<const uint4 *>(src + i);
    }
}

/*
 * Shared memory plan of the DFA kernels (offsets from the dynamic smem base):
 *   [0, tab)            transition table
 *   [.., +fin)          fin[nstates]
 *   [.., +256)          byte-class map (class-compressed variant only)
 *   [bar_ofs, ...)      per-warp mbarriers of the TMA pipeline
 *   [stage_ofs, ...)    per-warp staging rings (k_dfa_lines only), 1 KB aligned
 */
struct dfa_smem_plan_t {
    size_t tab_bytes, fin_ofs, cls_ofs, bar_ofs, stage_ofs;
};
constexpr int MAX_WARPS = 32, MAX_STAGES = 8;

__host__ __device__ inline dfa_smem_plan_t dfa_smem_plan(uint32_t nstates, uint32_t nclasses, bool cls)
{
    dfa_smem_plan_t p;
    /* (the [state][byte] tables: room for rows padded to ROW260 bytes) */
    p.tab_bytes = align_up(cls ? (size_t) nstates * nclasses * 2 : (size_t) nstates * ROW260, 16);
    p.fin_ofs = p.tab_bytes;
    p.cls_ofs = p.fin_ofs + align_up(nstates, 16);
    p.bar_ofs = align_up(p.cls_ofs + (cls ? 256 : 0), 16);
    p.stage_ofs = align_up(p.bar_ofs + MAX_WARPS * MAX_STAGES * 8, 1024);
    return p;
}


/* ---- TMA (cp.async.bulk.tensor) + mbarrier helpers ------------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

/* one 2-D box {c0 .. c0+box0, c1 .. c1+box1} -> dst, completes on bar */
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int32_t c0, int32_t c1,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

/*
 * Per-warp tile pipeline (TMA, early stage release).  The tensor map describes
 * the corpus as a 2-D byte tensor {pitch, nrows}; the warp owns row groups
 * g = gw, gw + warps_total, ... (32 consecutive rows each; lane l consumes row
 * g*32 + l).  One elected lane asks the TMA unit for a {128 bytes x 32 rows} box
 * per tile, written with the hardware's 128-byte swizzle (so that "lane l reads
 * 16-byte chunk c of row l" is bank-conflict free) and signalled on a per-stage
 * mbarrier; rows past nrows and bytes past the pitch are zero-filled by the
 * hardware.  A lane copies its row of the landed tile into registers in two
 * 64-byte halves; as soon as the second half is in registers (half-way through
 * the tile's processing time) the stage buffer is handed back to the TMA unit.
 * The staging memory per row is therefore only STAGES x 128 bytes with STAGES =
 * 1 or 2, which lets 32-48 warps per SM be resident: the per-byte dependent
 * chain (PRMT -> LDS.U8, ~35 cycles) needs that many independent rows in
 * flight to keep the shared-memory pipe busy.
 *
 * Consumer interface:
 *   void begin(size_t group);      my row of `group` starts
 *   void chunk(const uint4 &v);    next 16 bytes of my row
 *   void byte(uint32_t b);         next single byte (ragged tail of a row)
 *   void end(size_t group);        my row of `group` is complete
 */
template <int STAGES, class Consumer>
__device__ __forceinline__ void tile_pipeline_tma_early(Consumer &cons, const CUtensorMap *tmap, size_t nrows,
    uint32_t rowlen, uint8_t *my_stage, uint64_t *my_bars, size_t gw, size_t warps_total)
{
    constexpr int TW = 128;
    constexpr int STAGE_BYTES = 32 * TW;
    const uint32_t lane = threadIdx.x & 31;
    const size_t ngroups = (nrows + 31) / 32;
    if (gw >= ngroups) {
        return;
    }
    const uint32_t my_groups = (uint32_t) ((ngroups - gw + warps_total - 1) / warps_total);
    const uint32_t ntiles = (rowlen + TW - 1) / TW;

    if (ntiles == 0) {
        for (uint32_t gi = 0; gi < my_groups; gi++) {
            cons.begin(gw + (size_t) gi * warps_total);
            cons.end(gw + (size_t) gi * warps_total);
        }
        return;
    }

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&my_bars[i], 1);
        }
        fence_mbar_init();
    }
    __syncwarp();

    const uint32_t total = my_groups * ntiles;
    uint32_t pk = 0, pgi = 0, pt = 0;
    auto produce = [&]() {
        if (pk < total) {
            if (lane == 0) {
                uint64_t *bar = &my_bars[pk % STAGES];
                /* the lanes' (generic-proxy) reads of this stage are ordered before the
                 * __syncwarp that precedes this call; the fence orders them before the
                 * TMA unit's (async-proxy) write into the same bytes */
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                mbar_arrive_expect_tx(bar, STAGE_BYTES);
                tma_load_2d(my_stage + (pk % STAGES) * STAGE_BYTES, tmap, (int32_t) (pt * TW),
                            (int32_t) ((gw + (size_t) pgi * warps_total) * 32), bar);
            }
            pk++;
            if (++pt == ntiles) {
                pt = 0;
                pgi++;
            }
        }
    };
#pragma unroll
    for (int i = 0; i < STAGES; i++) {
        produce();
    }

    uint32_t t = 0;
    size_t group = gw;
    const uint32_t swz = lane & 7;
    cons.begin(group);

    for (uint32_t k = 0; k < total; k++) {
        mbar_wait(&my_bars[k % STAGES], (k / STAGES) & 1);
        const uint8_t *row = my_stage + (k % STAGES) * STAGE_BYTES + lane * TW;
        const uint32_t left = rowlen - t * TW;

        uint4 a[4], b[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            a[c] = *reinterpret_cast<const uint4 *>(row + ((c ^ swz) << 4));
        }
        if (left >= (uint32_t) TW) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                cons.chunk(a[c]);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                b[c] = *reinterpret_cast<const uint4 *>(row + (((c + 4) ^ swz) << 4));
            }
            __syncwarp();
            produce();
#pragma unroll
            for (int c = 0; c < 4; c++) {
                cons.chunk(b[c]);
            }
        } else {
            /* ragged last tile of a row: byte loop straight from the stage */
            for (uint32_t i = 0; i < left; i++) {
                cons.byte(row[(((i >> 4) ^ swz) << 4) | (i & 15)]);
            }
            __syncwarp();
            produce();
        }

        if (++t == ntiles) {
            cons.end(group);
            t = 0;
            group += warps_total;
            if (k + 1 < total) {
                cons.begin(group);
            }
        }
    }
}

int num_sms();
/* host: 2-D byte tensor {pitch, nrows}, box {128, 32}, 128-byte swizzle */
cudaError_t make_row_tensor_map(CUtensorMap *map, const uint8_t *buf, size_t nrows, size_t pitch, int tw);
constexpr size_t SMEM_LIMIT = 227 * 1024;

}  // namespace sre_dev

#endif
