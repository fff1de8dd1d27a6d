/*
 * sre_device_common.cuh -- device helpers shared by the Thompson kernels:
 * cp.async wrappers, DFA step functors, the shared-memory plan of the DFA
 * tables, and the per-warp tile pipeline that stages 32 rows x TW bytes through
 * shared memory (rows = lines for k_dfa_lines, stream pieces for
 * k_stream_pieces).
 */
#ifndef SRE_DEVICE_COMMON_CUH
#define SRE_DEVICE_COMMON_CUH

#include "sre_kernels.cuh"

namespace sre_dev {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem) : "memory");
}

__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory");
}

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
 *   [stage_ofs, ...)    per-warp staging rings (k_dfa_lines only), 1 KB aligned
 */
struct dfa_smem_plan_t {
    size_t tab_bytes, fin_ofs, cls_ofs, stage_ofs;
};

__host__ __device__ inline dfa_smem_plan_t dfa_smem_plan(uint32_t nstates, uint32_t nclasses, bool cls)
{
    dfa_smem_plan_t p;
    p.tab_bytes = align_up(cls ? (size_t) nstates * nclasses * 2 : (size_t) nstates * 256, 16);
    p.fin_ofs = p.tab_bytes;
    p.cls_ofs = p.fin_ofs + align_up(nstates, 16);
    p.stage_ofs = align_up(p.cls_ofs + (cls ? 256 : 0), 1024);
    return p;
}


template <int TW>
__device__ __forceinline__ uint32_t swizzle(uint32_t row)
{
    /* 16-byte chunks per row; rows that share a 128-byte bank window get
     * distinct XOR keys so that "lane l reads chunk c of row l" touches every
     * bank once per quarter-warp */
    constexpr int CPR = TW / 16;
    if (CPR >= 8) return row & 7;
    if (CPR == 4) return (row >> 1) & 3;
    if (CPR == 2) return (row >> 2) & 1;
    return 0;
}

/*
 * Per-warp tile pipeline.  The warp owns row groups g = gw, gw + warps_total,
 * ... (32 consecutive rows each; lane l consumes row g*32 + l).  Every row is
 * rowlen bytes at buf + row*pitch (16-byte aligned).  Tiles of TW bytes per row
 * travel through a STAGES-deep ring of cp.async groups that runs ahead across
 * group boundaries, so the copy engine never drains between groups.
 *
 * Consumer interface:
 *   void begin();                  a new row starts
 *   void chunk(const uint4 &v);    next 16 bytes of my row
 *   void byte(uint32_t b);         next single byte (ragged tail of a row)
 *   void end(size_t group);        my row of `group` is complete
 */
template <int TW, int STAGES, class Consumer>
__device__ __forceinline__ void tile_pipeline(Consumer &cons, const uint8_t *__restrict__ buf,
    size_t nrows, size_t pitch, uint32_t rowlen, uint8_t *my_stage, size_t gw, size_t warps_total)
{
    constexpr int CPR = TW / 16;
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
            cons.begin();
            cons.end(gw + (size_t) gi * warps_total);
        }
        return;
    }

    const uint32_t total = my_groups * ntiles;

    auto issue = [&](uint32_t k) {
        const uint32_t gi = k / ntiles, t = k - gi * ntiles;
        const size_t group = gw + (size_t) gi * warps_total;
        uint8_t *dst = my_stage + (k % STAGES) * STAGE_BYTES;
        const uint32_t tile_off = t * TW;
        const uint32_t left = rowlen - tile_off;
        const uint32_t valid = left >= (uint32_t) TW ? (uint32_t) TW : (left + 15) & ~15u;
#pragma unroll
        for (int j = 0; j < CPR; j++) {
            const uint32_t q = lane + 32 * j, row = q / CPR, c = q % CPR;
            size_t r = group * 32 + row;
            if (r >= nrows) {
                r = nrows - 1;
            }
            if (c * 16 < valid) {
                cp_async16(dst + row * TW + ((c ^ swizzle<TW>(row)) << 4),
                           buf + r * pitch + tile_off + c * 16);
            }
        }
    };

    uint32_t issued = 0;
#pragma unroll
    for (int i = 0; i < STAGES - 1; i++) {
        if (issued < total) {
            issue(issued);
        }
        cp_async_commit();
        issued++;
    }

    uint32_t t = 0;
    size_t group = gw;
    const uint32_t swz = swizzle<TW>(lane);
    cons.begin();

    for (uint32_t k = 0; k < total; k++) {
        if (issued < total) {
            issue(issued);
        }
        cp_async_commit();
        issued++;
        cp_async_wait<STAGES - 1>();
        __syncwarp();

        const uint8_t *row = my_stage + (k % STAGES) * STAGE_BYTES + lane * TW;
        const uint32_t left = rowlen - t * TW;
        if (left >= (uint32_t) TW) {
#pragma unroll
            for (int c = 0; c < CPR; c++) {
                cons.chunk(*reinterpret_cast<const uint4 *>(row + ((c ^ swz) << 4)));
            }
        } else {
            for (uint32_t i = 0; i < left; i++) {
                cons.byte(row[(((i >> 4) ^ swz) << 4) | (i & 15)]);
            }
        }
        __syncwarp();

        if (++t == ntiles) {
            cons.end(group);
            t = 0;
            group += warps_total;
            cons.begin();
        }
    }
    cp_async_wait<0>();
}

int num_sms();
constexpr size_t SMEM_LIMIT = 227 * 1024;

}  // namespace sre_dev

#endif
