/*
 * sre_stream.cu -- chunk-parallel Thompson matching of ONE long stream.
 *
 * What it replaces: a sequence of sre_vm_thompson_exec(ctx, chunk, n, eof)
 * calls with SRE_AGAIN state carry (reference sre_vm_thompson.c:63-270, ctx
 * fields sre_vm_thompson.h:30-41).  The carried state of the determinised
 * program is one DFA state, so a piece of the stream is a function
 * f: state -> state and composition is associative.  A function on thousands
 * of states is too big to compute per piece, so a piece is summarised on the
 * few states it can actually be entered in: the IMAGE of all states under the
 * WINDOW bytes in front of the piece, which the image automaton
 * (../lower/sre_image.h) yields with one table walk.  The true entry state is
 * always among these candidates, so nothing below depends on a guess:
 *
 *   k_stream_pieces   one CUDA thread per PIECE-byte piece (same TMA tile
 *                     pipeline as k_dfa_lines_tma_early, rows = pieces): walks
 *                     the image automaton over the window in front of the
 *                     piece, runs the piece from each candidate (<= K) until
 *                     they have met in one state or been absorbed by ACC, then
 *                     the single-state inner loop; writes the piece's record
 *                     {candidate -> exit state} (32 bytes).  A window that
 *                     leaves more than K candidates marks the piece
 *                     "unresolved".
 *   k_stream_tail     the ragged end (< PIECE bytes): one warp, a sub-span per
 *                     lane with its own window, shuffle-tree composition.
 *   k_stream_compose  one warp composes FAN = 256 consecutive records (8 per
 *                     lane serially, then a 5-round shuffle tree) into one
 *                     record of the next level.  Composition looks the exit
 *                     state of the left record up among the candidates of the
 *                     right one; "unresolved" is absorbing.
 *   k_stream_fix      for the first unresolved piece (found by walking down the
 *                     levels): the state the stream enters it in (walking the
 *                     levels left of it), then the piece itself run from that
 *                     state; its record becomes {entry -> exit} and the
 *                     ancestors are recomposed.  Repeats until none is left.
 *   k_stream_descend  one warp walks back down from the true entry state along
 *                     the only path that matters: at each level the child in
 *                     whose span ACC is first entered (= the chunk in which the
 *                     reference's loop returns SRE_OK); also yields the state
 *                     after the whole stream.
 *   k_stream_locate   one warp re-runs that piece to get the exact byte offset.
 */
#include <cstring>
#include <type_traits>

#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

constexpr uint32_t PIECE = 4096;        /* bytes per level-0 piece                 */
constexpr uint32_t FAN = 256;           /* records composed per warp               */
constexpr uint32_t PER_LANE = FAN / 32;
constexpr int      K = SRE_STREAM_K;    /* candidates per record                   */
constexpr uint32_t WINDOW = 64;         /* bytes in front of a piece walked first  */
constexpr uint32_t WINDOW2 = 1024;      /* ... and when that window stays wide     */
constexpr uint32_t ACC = 1;             /* absorbing state of every lowered DFA    */
constexpr uint32_t NONE = 0xffffu;      /* empty slot / "not among the candidates" */
constexpr uint32_t UNRESOLVED = 0xfffeu;
constexpr uint32_t IDENTITY = 0xfffdu;

/* ---- a record: K candidate entry states and the states they lead to ---------------- */

struct rec_t {
    uint32_t w[K];      /* u16 pairs: w[0 .. K/2) candidates, w[K/2 .. K) exits */

    __host__ __device__ __forceinline__ uint32_t cand(int j) const { return (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu; }
    __host__ __device__ __forceinline__ uint32_t exit(int j) const
    {
        return (w[K / 2 + (j >> 1)] >> ((j & 1) * 16)) & 0xffffu;
    }
    __host__ __device__ __forceinline__ void set(int j, uint32_t c, uint32_t e)
    {
        const uint32_t sh = (j & 1) * 16, m = ~(0xffffu << sh);
        w[j >> 1] = (w[j >> 1] & m) | (c << sh);
        w[K / 2 + (j >> 1)] = (w[K / 2 + (j >> 1)] & m) | (e << sh);
    }
    __host__ __device__ __forceinline__ void clear(uint32_t marker)
    {
#pragma unroll
        for (int i = 0; i < K; i++) {
            w[i] = 0xffffffffu;
        }
        w[0] = 0xffff0000u | marker;
    }
    __host__ __device__ __forceinline__ void identity() { clear(IDENTITY); }
    __host__ __device__ __forceinline__ void unresolved() { clear(UNRESOLVED); }
    __host__ __device__ __forceinline__ bool is_identity() const { return (w[0] & 0xffffu) == IDENTITY; }
    __host__ __device__ __forceinline__ bool is_unresolved() const { return (w[0] & 0xffffu) == UNRESOLVED; }

    /* state reached from s; NONE when s is not among the candidates */
    __host__ __device__ __forceinline__ uint32_t at(uint32_t s) const
    {
        if (s == ACC || is_identity()) {
            return s;
        }
        uint32_t r = NONE;
#pragma unroll
        for (int j = 0; j < K; j++) {
            r = cand(j) == s ? exit(j) : r;
        }
        return r;
    }
    /* this = g o this (apply this first, then g) */
    __host__ __device__ __forceinline__ void then(const rec_t &g)
    {
        if (is_unresolved() || g.is_identity()) {
            return;
        }
        if (g.is_unresolved() || is_identity()) {
            *this = g;
            return;
        }
        bool bad = false;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const uint32_t c = cand(j);
            if (c != NONE) {
                const uint32_t e = g.at(exit(j));
                bad |= e == NONE;
                set(j, c, e);
            }
        }
        if (bad) {
            /* cannot happen (the right record's candidates contain every state the
             * stream can be in there); were it to, the piece is redone serially */
            unresolved();
        }
    }
#ifdef __CUDACC__
    __device__ __forceinline__ void load(const uint8_t *p)
    {
        const uint4 a = *reinterpret_cast<const uint4 *>(p), b = *reinterpret_cast<const uint4 *>(p + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
    __device__ __forceinline__ void store(uint8_t *p) const
    {
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4 *>(p + 16) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    __device__ __forceinline__ rec_t shfl_down(uint32_t delta) const
    {
        rec_t r;
#pragma unroll
        for (int i = 0; i < K; i++) {
            r.w[i] = __shfl_down_sync(FULL, w[i], delta);
        }
        return r;
    }
#endif
};
static_assert(K == 8 && sizeof(rec_t) == SRE_STREAM_FN_BYTES, "record layout");

/* compose the records held by the 32 lanes in lane order; lane 0 gets the result */
__device__ __forceinline__ void warp_compose(rec_t &f)
{
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (uint32_t s = 1; s < 32; s <<= 1) {
        const rec_t g = f.shfl_down(s);
        if ((lane & (2 * s - 1)) == 0) {
            f.then(g);
        }
    }
}

/* ---- candidates of a position ------------------------------------------------------ */

/* what every stream kernel knows about the stream */
struct stream_src_t {
    const uint8_t  *buf;        /* this call's part of the stream (16-byte aligned)         */
    size_t          len;
    const uint8_t  *halo;       /* SRE_STREAM_HALO bytes that precede buf, or NULL          */
    uint32_t        entry;      /* state the stream enters buf in, or SRE_STREAM_UNKNOWN    */
};

/* one step of the image automaton (clsmap: the DFA's byte classes) */
__device__ __forceinline__ uint32_t image_step(const sre_dev_image_t &img, const uint8_t *clsmap, uint32_t u,
                                               uint32_t b)
{
    return __ldg(img.trans + u * img.nclasses + clsmap[b]);
}

/* U-state of the window buf[from, to) walked from TOP (byte loads) */
__device__ __forceinline__ uint32_t image_walk(const sre_dev_image_t &img, const uint8_t *clsmap, const uint8_t *buf,
                                               size_t from, size_t to)
{
    uint32_t u = 0;
    for (size_t i = from; i < to; i++) {
        u = image_step(img, clsmap, u, __ldg(buf + i));
    }
    return u;
}

/*
 * Which candidate set position p of the stream part has (p a multiple of 16 when
 * VEC): an image-automaton state, or ENTRY_U for "exactly the known entry
 * state"; n = how many candidates, -1 when more than K remain (wide).
 */
constexpr uint32_t ENTRY_U = 0xffffffffu;

template <bool VEC>
__device__ __forceinline__ int candidates_id(const sre_dev_image_t &img, const uint8_t *clsmap,
                                             const stream_src_t &src, size_t p, uint32_t &u)
{
    u = 0;
    if (p == 0 && src.entry != SRE_STREAM_UNKNOWN) {
        u = ENTRY_U;
        return src.entry == ACC ? 0 : 1;
    }
    if (p == 0) {
        if (src.halo == nullptr) {
            return -1;
        }
        u = image_walk(img, clsmap, src.halo, 0, SRE_STREAM_HALO);
        const uint32_t n = __ldg(img.ncand + u);
        return n == 0xff ? -1 : (int) n;
    }
    const size_t from = p > WINDOW ? p - WINDOW : 0;
    if (VEC && p >= WINDOW) {
#pragma unroll 1
        for (size_t i = from; i < p; i += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src.buf + i));
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int q = 0; q < 16; q++) {
                u = image_step(img, clsmap, u, (w[q >> 2] >> ((q & 3) * 8)) & 0xff);
            }
        }
    } else {
        u = image_walk(img, clsmap, src.buf, from, p);
    }
    uint32_t n = __ldg(img.ncand + u);
    if (n == 0xff && from > 0) {
        /* a window that never leaves TOP or a wide set: try a longer one */
        u = image_walk(img, clsmap, src.buf, p > WINDOW2 ? p - WINDOW2 : 0, p);
        n = __ldg(img.ncand + u);
    }
    return n == 0xff ? -1 : (int) n;
}

/* the n candidates of set u; unused slots repeat the first one (ACC when there
 * is none), so that running "all K" of them is harmless */
__device__ __forceinline__ void candidates_load(const sre_dev_image_t &img, const stream_src_t &src, uint32_t u,
                                                int n, uint32_t (&cand)[K])
{
    if (u == ENTRY_U) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            cand[j] = src.entry;
        }
        return;
    }
    const uint4 c = __ldg(reinterpret_cast<const uint4 *>(img.cand + (size_t) u * K));
    cand[0] = c.x & 0xffff; cand[1] = c.x >> 16; cand[2] = c.y & 0xffff; cand[3] = c.y >> 16;
    cand[4] = c.z & 0xffff; cand[5] = c.z >> 16; cand[6] = c.w & 0xffff; cand[7] = c.w >> 16;
    const uint32_t fill = n > 0 ? cand[0] : ACC;
#pragma unroll
    for (int j = 0; j < K; j++) {
        cand[j] = j < n ? cand[j] : fill;
    }
}

/* record from the candidates and where each of them ended up */
__device__ __forceinline__ void make_record(rec_t &r, int n, const uint32_t (&cand)[K], const uint32_t (&cur)[K])
{
    if (n < 0) {
        r.unresolved();
        return;
    }
    r.clear(NONE);
#pragma unroll
    for (int j = 0; j < K; j++) {
        if (j < n) {
            r.set(j, cand[j], cur[j]);
        }
    }
}

/* ---- level 0: pieces ------------------------------------------------------------------ */

/* class-compressed table through the read-only path (tables beyond shared memory) */
struct stepcls_g_t {
    const uint16_t *tab;
    const uint8_t  *cls;        /* shared memory */
    uint32_t        ncls;
    __device__ __forceinline__ uint32_t byte(uint32_t s, uint32_t b) const { return __ldg(tab + s * ncls + cls[b]); }
    __device__ __forceinline__ uint32_t word(uint32_t s, uint32_t w) const
    {
        s = byte(s, w & 0xff);
        s = byte(s, (w >> 8) & 0xff);
        s = byte(s, (w >> 16) & 0xff);
        s = byte(s, w >> 24);
        return s;
    }
};

template <class Step>
struct piece_consumer_t {
    Step                step;
    sre_dev_image_t     img;
    const uint8_t      *clsmap;     /* shared memory */
    stream_src_t        src;
    uint8_t            *fn;         /* level-0 records */
    size_t              npieces;
    uint32_t            cur[K];     /* where each candidate is now               */
    uint32_t            s;          /* the single live state once converged      */
    uint32_t            accmask;    /* candidates already absorbed by ACC        */
    uint32_t            u;          /* which candidate set (candidates_id)       */
    int                 n;          /* candidates, -1: unresolved                */
    bool                conv;

    __device__ __forceinline__ void begin(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        conv = true;
        s = ACC;
        accmask = 0;
        n = 0;
        if (piece >= npieces) {
            return;
        }
        n = candidates_id<true>(img, clsmap, src, piece * PIECE, u);
        if (n < 0) {
            return;
        }
        candidates_load(img, src, u, n, cur);
        if (n > 1) {
            conv = false;
            check();
        } else if (n == 1) {
            s = cur[0];
        }
    }

    /* converged = every candidate has been absorbed by ACC or has reached one common state */
    __device__ __forceinline__ void check()
    {
        uint32_t ref = ACC, mask = 0;
        bool all = true;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const uint32_t v = cur[j];
            if (v == ACC) {
                mask |= 1u << j;
            } else if (ref == ACC) {
                ref = v;
            } else if (v != ref) {
                all = false;
            }
        }
        if (all) {
            conv = true;
            s = ref;
            accmask = mask;
        }
    }

    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        if (conv) {
            s = step.word(step.word(step.word(step.word(s, v.x), v.y), v.z), v.w);
            return;
        }
#pragma unroll
        for (int j = 0; j < K; j++) {
            cur[j] = step.word(step.word(step.word(step.word(cur[j], v.x), v.y), v.z), v.w);
        }
        check();
    }

    __device__ __forceinline__ void byte(uint32_t b)
    {
        if (conv) {
            s = step.byte(s, b);
            return;
        }
#pragma unroll
        for (int j = 0; j < K; j++) {
            cur[j] = step.byte(cur[j], b);
        }
    }

    __device__ __forceinline__ void end(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        if (piece >= npieces) {
            return;
        }
        rec_t r;
        if (n < 0) {
            r.unresolved();
        } else {
            /* the candidates again (they were overwritten by the run) */
            uint32_t cand[K];
            candidates_load(img, src, u, n, cand);
            if (conv) {
#pragma unroll
                for (int j = 0; j < K; j++) {
                    cur[j] = ((accmask >> j) & 1) ? ACC : s;
                }
            }
            make_record(r, n, cand, cur);
        }
        r.store(fn + piece * sizeof(rec_t));
    }
};

/* shared memory: [table][clsmap 256][barriers][stages]; TAB: 0 = [state][byte] rows of 260 bytes,
 * 1 = class table in shared memory, 2 = class table through the read-only path */
struct stream_smem_plan_t {
    size_t cls_ofs, bar_ofs, stage_ofs;
};
__host__ __device__ inline stream_smem_plan_t stream_smem_plan(const sre_dev_dfa_t &dfa, int tab)
{
    stream_smem_plan_t p;
    const size_t t = tab == 0 ? (size_t) (dfa.nstates + 1) * ROW260 : tab == 1 ? (size_t) dfa.nstates * dfa.nclasses * 2 : 0;
    p.cls_ofs = align_up(t, 16);
    p.bar_ofs = p.cls_ofs + 256;
    p.stage_ofs = align_up(p.bar_ofs + MAX_WARPS * MAX_STAGES * 8, 1024);
    return p;
}

template <int TAB>
__global__ void __launch_bounds__(1024, 1)
k_stream_pieces(sre_dev_dfa_t dfa, sre_dev_image_t img, const __grid_constant__ CUtensorMap tmap, stream_src_t src,
                size_t npieces, uint8_t *fn)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const stream_smem_plan_t plan = stream_smem_plan(dfa, TAB);
    if (TAB == 0) {
        load_table260(smem, dfa.t256, dfa.nstates);
    } else if (TAB == 1) {
        load_table(smem, reinterpret_cast<const uint8_t *>(dfa.tcls), align_up((size_t) dfa.nstates * dfa.nclasses * 2, 16));
    }
    load_table(smem + plan.cls_ofs, dfa.clsmap, 256);
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    using step_t = typename std::conditional<TAB == 0, step260_t,
                   typename std::conditional<TAB == 1, stepcls_t, stepcls_g_t>::type>::type;
    piece_consumer_t<step_t> cons;
    if constexpr (TAB == 0) {
        cons.step.tab_s = (uint32_t) __cvta_generic_to_shared(smem);
    } else if constexpr (TAB == 1) {
        cons.step.tab = reinterpret_cast<const uint16_t *>(smem);
        cons.step.cls = smem + plan.cls_ofs;
        cons.step.ncls = dfa.nclasses;
    } else {
        cons.step.tab = dfa.tcls;
        cons.step.cls = smem + plan.cls_ofs;
        cons.step.ncls = dfa.nclasses;
    }
    cons.img = img;
    cons.clsmap = smem + plan.cls_ofs;
    cons.src = src;
    cons.fn = fn;
    cons.npieces = npieces;

    /* rows = pieces: the stream is a {PIECE, npieces} byte tensor */
    tile_pipeline_tma_early<1>(cons, &tmap, npieces, PIECE,
                               smem + plan.stage_ofs + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + plan.bar_ofs) + warp * MAX_STAGES,
                               /* block-major interleave: a partial last round of
                                * groups is spread over all SMs, not over the first few */
                               (size_t) warp * gridDim.x + blockIdx.x,
                               (size_t) gridDim.x * warps_per_block);
}

/* ---- one-warp kernels: global tables ---------------------------------------------------- */

__device__ __forceinline__ uint32_t dfa_step_g(const sre_dev_dfa_t &dfa, uint32_t s, uint32_t b)
{
    return __ldg(dfa.tcls + s * dfa.nclasses + __ldg(dfa.clsmap + b));
}

/* record of the span [b, e) of the stream part, candidates from the window in front of b */
__device__ __forceinline__ void span_record(const sre_dev_dfa_t &dfa, const sre_dev_image_t &img,
                                            const stream_src_t &src, size_t b, size_t e, rec_t &r)
{
    if (b >= e && !(b == 0 && src.len == 0)) {
        r.identity();
        return;
    }
    uint32_t cand[K], cur[K], u;
    const int n = candidates_id<false>(img, dfa.clsmap, src, b, u);
    if (n < 0) {
        r.unresolved();
        return;
    }
    candidates_load(img, src, u, n, cand);
#pragma unroll
    for (int j = 0; j < K; j++) {
        cur[j] = cand[j];
    }
    for (size_t i = b; i < e; i++) {
        const uint32_t c = __ldg(dfa.clsmap + __ldg(src.buf + i));
#pragma unroll
        for (int j = 0; j < K; j++) {
            cur[j] = __ldg(dfa.tcls + cur[j] * dfa.nclasses + c);
        }
    }
    make_record(r, n, cand, cur);
}

/* the 32 lanes' sub-spans of [begin, end): lane l covers [sb, se) */
__device__ __forceinline__ void lane_span(size_t begin, size_t end, size_t &sb, size_t &se)
{
    const uint32_t lane = threadIdx.x & 31;
    /* 16-byte aligned cuts keep most windows on vector-friendly boundaries */
    const size_t sub = ((end - begin + 31) / 32 + 15) & ~(size_t) 15;
    sb = begin + lane * sub < end ? begin + lane * sub : end;
    se = sb + sub < end ? sb + sub : end;
}

/* the ragged tail [begin, len) (< PIECE bytes; also the whole of an empty part) */
__global__ void __launch_bounds__(32)
k_stream_tail(sre_dev_dfa_t dfa, sre_dev_image_t img, stream_src_t src, size_t begin, uint8_t *out)
{
    size_t sb, se;
    lane_span(begin, src.len, sb, se);
    rec_t f;
    if (src.len == begin) {
        /* nothing to read: the identity, except for an empty part of a stream
         * whose entry state is wanted as a one-candidate record */
        f.identity();
    } else {
        span_record(dfa, img, src, sb, se, f);
    }
    warp_compose(f);
    if ((threadIdx.x & 31) == 0) {
        f.store(out);
    }
}

/* ---- upper levels --------------------------------------------------------------------------- */

/* lane's local composition of its PER_LANE consecutive records */
__device__ __forceinline__ void lane_compose(const uint8_t *in, size_t first, size_t n_in, size_t limit, rec_t &f)
{
    f.identity();
#pragma unroll
    for (uint32_t k = 0; k < PER_LANE; k++) {
        if (first + k < n_in && first + k < limit) {
            rec_t g;
            g.load(in + (first + k) * sizeof(rec_t));
            f.then(g);
        }
    }
}

/* out[w] = in[w*FAN + FAN-1] o ... o in[w*FAN], one warp per output; only_out (device, may be
 * NULL): recompose just the ancestor of piece *only_out at this level (span = pieces per output) */
__global__ void __launch_bounds__(256)
k_stream_compose(const uint8_t *__restrict__ in, size_t n_in, uint8_t *__restrict__ out, size_t n_out,
                 const unsigned long long *only_out, unsigned long long span)
{
    size_t w = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (only_out != nullptr) {
        if (w != 0 || *only_out == ~0ull) {
            return;
        }
        w = (size_t) (*only_out / span);
    }
    if (w >= n_out) {
        return;
    }
    rec_t f;
    lane_compose(in, w * FAN + lane * PER_LANE, n_in, ~(size_t) 0, f);
    warp_compose(f);
    if (lane == 0) {
        f.store(out + w * sizeof(rec_t));
    }
}

struct stream_levels_t {
    uint8_t       *fn[4];
    size_t         count[4];
    int            top;
};

/* state after the first `upto` children of `parent` at level l, entered in s_in (whole warp);
 * also returns in `mine` the state in which this lane's first child is entered */
__device__ __forceinline__ uint32_t walk_children(const stream_levels_t &lv, int l, size_t parent, size_t upto,
                                                  uint32_t s_in, uint32_t &mine)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t first = parent * FAN + lane * PER_LANE;
    rec_t total;
    lane_compose(lv.fn[l], first, lv.count[l], upto, total);
    uint32_t s = s_in;
    mine = 0;
#pragma unroll 4
    for (uint32_t k = 0; k < 32; k++) {
        if (lane == k) {
            mine = s;
        }
        s = __shfl_sync(FULL, s == NONE ? NONE : total.at(s), k);
    }
    return s;
}

/*
 * The first unresolved piece, if any: found by walking down the levels (an
 * unresolved record makes every ancestor unresolved), its entry state computed
 * on the way (the siblings left of it are all resolved), the piece run from
 * that state by lane 0, its record rewritten as {entry -> exit}.  *fixed = the
 * piece, or ~0 when every piece is resolved (then nothing is touched).
 */
__global__ void __launch_bounds__(32)
k_stream_fix(sre_dev_dfa_t dfa, stream_src_t src, stream_levels_t lv, uint32_t root_state,
             unsigned long long *fixed)
{
    const uint32_t lane = threadIdx.x;
    size_t parent = 0;
    uint32_t s_in = root_state;
    for (int l = lv.top; l >= 0; l--) {
        const size_t n = lv.count[l], first = parent * FAN + lane * PER_LANE;
        size_t bad = ~(size_t) 0;
#pragma unroll
        for (uint32_t k = 0; k < PER_LANE; k++) {
            const size_t j = first + k;
            if (j < n && j < (parent + 1) * FAN && bad == ~(size_t) 0) {
                rec_t g;
                g.load(lv.fn[l] + j * sizeof(rec_t));
                if (g.is_unresolved()) {
                    bad = j;
                }
            }
        }
        const uint32_t who = __ffs(__ballot_sync(FULL, bad != ~(size_t) 0));
        if (who == 0) {
            /* only possible at the top: everything is resolved */
            if (lane == 0) {
                *fixed = ~0ull;
            }
            return;
        }
        const size_t child = (size_t) __shfl_sync(FULL, (unsigned long long) bad, who - 1);
        uint32_t mine;
        s_in = walk_children(lv, l, parent, child, s_in, mine);
        parent = child;
    }
    /* parent = the piece, s_in = the state it is entered in */
    if (lane == 0) {
        const size_t b = parent * PIECE, e = b + PIECE < src.len ? b + PIECE : src.len;
        uint32_t s = s_in;
        for (size_t i = b; i < e && s != ACC && s != NONE; i++) {
            s = dfa_step_g(dfa, s, __ldg(src.buf + i));
        }
        rec_t r;
        r.clear(NONE);
        if (s_in != ACC && s_in != NONE) {
            r.set(0, s_in, s);
        }
        r.store(lv.fn[0] + parent * sizeof(rec_t));
        *fixed = (unsigned long long) parent;
    }
}

/*
 * The walk down needs one path only: at every level, the child of the current
 * parent in whose span a non-ACC state is taken into ACC (ACC is absorbing, so
 * that is the child in which the match is first seen).  Outputs: the state
 * after the whole stream part, the first piece in which ACC is entered (or ~0)
 * and the state that piece is entered in (for k_stream_locate).
 */
__global__ void __launch_bounds__(32)
k_stream_descend(stream_levels_t lv, uint32_t root_state, unsigned long long *first_acc, uint32_t *out)
{
    const uint32_t lane = threadIdx.x;
    size_t parent = 0;
    uint32_t s_in = root_state;
    for (int l = lv.top; l >= 0; l--) {
        const size_t n = lv.count[l], first = parent * FAN + lane * PER_LANE;
        uint32_t mine;
        uint32_t s = walk_children(lv, l, parent, (parent + 1) * FAN, s_in, mine);
        if (l == lv.top && lane == 0) {
            out[0] = s == NONE ? 0xffffffffu : s;       /* exit state */
        }
        s = mine;
        unsigned long long hit = ~0ull;
        uint32_t hit_entry = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER_LANE; k++) {
            const size_t j = first + k;
            if (j < n && j < (parent + 1) * FAN && s != NONE) {
                rec_t g;
                g.load(lv.fn[l] + j * sizeof(rec_t));
                const uint32_t nx = g.at(s);
                if (nx == ACC && s != ACC && hit == ~0ull) {
                    hit = j;
                    hit_entry = s;
                }
                s = nx;
            }
        }
        const uint32_t who = __ffs(__ballot_sync(FULL, hit != ~0ull));
        if (who == 0) {
            if (lane == 0) {
                *first_acc = ~0ull;
            }
            return;
        }
        parent = (size_t) __shfl_sync(FULL, hit, who - 1);
        s_in = __shfl_sync(FULL, hit_entry, who - 1);
    }
    if (lane == 0) {
        *first_acc = (unsigned long long) parent;
        out[1] = s_in;
    }
}

/* exact offset of the byte whose step enters ACC inside piece *first_acc (entered in state[1]) */
__global__ void __launch_bounds__(32)
k_stream_locate(sre_dev_dfa_t dfa, sre_dev_image_t img, stream_src_t src, const unsigned long long *first_acc,
                const uint32_t *state, long long *match_offset)
{
    const uint32_t lane = threadIdx.x;
    const unsigned long long piece = *first_acc;
    if (piece == ~0ull) {
        if (lane == 0) {
            *match_offset = -1;
        }
        return;
    }
    const size_t start = (size_t) piece * PIECE, end = start + PIECE < src.len ? start + PIECE : src.len;
    size_t sb, se;
    lane_span(start, end, sb, se);
    rec_t f;
    if (lane == 0) {
        /* the piece's own entry state is known */
        stream_src_t one = src;
        one.buf = src.buf + start;
        one.len = se - start;
        one.halo = nullptr;
        one.entry = state[1];
        span_record(dfa, img, one, 0, se - sb, f);
    } else {
        span_record(dfa, img, src, sb, se, f);
    }
    const bool serial = __any_sync(FULL, f.is_unresolved());
    uint32_t s = state[1], mine = 0;
    if (!serial) {
        for (uint32_t l = 0; l < 32; l++) {
            if (lane == l) {
                mine = s;
            }
            s = __shfl_sync(FULL, s == NONE ? NONE : f.at(s), l);
        }
    }
    /* the first lane whose sub-span takes a non-ACC state into ACC */
    const bool hit = !serial && mine != ACC && mine != NONE && f.at(mine) == ACC;
    const uint32_t who = serial ? 1 : __ffs(__ballot_sync(FULL, hit));
    if (who == 0) {
        if (lane == 0) {
            *match_offset = -1;     /* cannot happen */
        }
        return;
    }
    if (lane == who - 1) {
        s = serial ? state[1] : mine;
        const size_t b = serial ? start : sb, e = serial ? end : se;
        long long at = -1;
        for (size_t i = b; i < e; i++) {
            s = dfa_step_g(dfa, s, __ldg(src.buf + i));
            if (s == ACC) {
                at = (long long) i;
                break;
            }
        }
        *match_offset = at;
    }
}

/* state in which `piece` is entered when the stream part is entered in root_state (whole warp;
 * every record left of the piece must be resolved) */
__device__ __forceinline__ uint32_t state_before_piece(const stream_levels_t &lv, uint32_t root_state, size_t piece)
{
    size_t parent = 0, span = 1;
    for (int l = 0; l < lv.top; l++) {
        span *= FAN;
    }
    uint32_t s = root_state;
    for (int l = lv.top; l >= 0; l--, span /= FAN) {
        const size_t child = piece / span;      /* the ancestor of the piece at level l */
        uint32_t mine;
        s = walk_children(lv, l, parent, child, s, mine);
        parent = child;
    }
    return s;
}

/*
 * Restart point for a leftmost-first (Pike) search of a long buffer whose match the scan has
 * located: the last position p0 < limit after a byte that only the ".*?" thread consumed (the
 * restart flag of the hint table, see k_dfa_lines_hint) -- no thread alive at p0 started before
 * it, so the Pike VM may begin there.  Walks back piece by piece from the one that holds
 * limit - 1; gives 0 after max_back pieces without a flag (always a valid place to begin).
 */
__global__ void __launch_bounds__(32)
k_stream_restart(sre_dev_dfa_t dfa, stream_src_t src, stream_levels_t lv, uint32_t root_state, size_t limit,
                 uint32_t max_back, long long *out)
{
    const uint32_t lane = threadIdx.x;
    if (limit == 0 || dfa.hcls == nullptr) {
        if (lane == 0) {
            *out = 0;
        }
        return;
    }
    size_t piece = (limit - 1) / PIECE;
    for (uint32_t back = 0;; back++) {
        const uint32_t s_in = state_before_piece(lv, root_state, piece);
        long long found = -1;
        if (lane == 0 && s_in != NONE) {
            const size_t b = piece * PIECE, e = b + PIECE < limit ? b + PIECE : limit;
            uint32_t s = s_in;
            for (size_t i = b; i < e && s != ACC; i++) {
                const uint32_t t = __ldg(dfa.hcls + s * dfa.hncls + __ldg(dfa.hclsmap + __ldg(src.buf + i)));
                s = t & 0x7fffu;
                if (t & 0x8000u) {
                    found = (long long) i + 1;
                }
            }
        }
        found = __shfl_sync(FULL, found, 0);
        if (found >= 0 || piece == 0 || back >= max_back || s_in == NONE) {
            if (lane == 0) {
                *out = found >= 0 ? found : 0;
            }
            return;
        }
        piece--;
    }
}

int top_level(const sre_stream_ws_t &ws)
{
    int top = 0;
    while (top < 3 && ws.count[top] > FAN) {
        top++;
    }
    return top;
}

stream_levels_t levels_of(const sre_stream_ws_t &ws)
{
    stream_levels_t lv;
    lv.top = top_level(ws);
    for (int l = 0; l < 4; l++) {
        lv.fn[l] = ws.fn[l];
        lv.count[l] = ws.count[l];
    }
    return lv;
}

template <int TAB>
cudaError_t launch_pieces(const sre_dev_dfa_t &dfa, const sre_dev_image_t &img, const stream_src_t &src,
    size_t npieces, uint8_t *fn, cudaStream_t stream)
{
    const stream_smem_plan_t plan = stream_smem_plan(dfa, TAB);
    const int warps = 32;
    const size_t smem = plan.stage_ofs + (size_t) warps * 32 * 128;
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, src.buf, npieces, PIECE, 128);
    if (err != cudaSuccess) {
        return err;
    }
    auto kern = k_stream_pieces<TAB>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (err != cudaSuccess) {
            return err;
        }
        smem_set = smem;
    }
    const size_t ngroups = (npieces + 31) / 32;
    size_t grid = (size_t) num_sms();
    const size_t need = (ngroups + warps - 1) / warps;
    if (grid > need) {
        grid = need;
    }
    kern<<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, img, tmap, src, npieces, fn);
    return cudaGetLastError();
}

}  // namespace

size_t sre_stream_piece_bytes(void) { return PIECE; }
uint32_t sre_stream_fan(void) { return FAN; }

/* host side of the record format (api/sre_cuda_api.cu composes the top level and the ranks' records) */
uint32_t sre_stream_fn_apply(const uint8_t *fn, uint32_t state)
{
    rec_t r;
    memcpy(r.w, fn, sizeof(r.w));
    if (state == ACC) {
        return ACC;             /* absorbing, whatever the part holds */
    }
    if (r.is_unresolved()) {
        return SRE_STREAM_UNKNOWN;
    }
    const uint32_t e = r.at(state);
    return e == NONE ? SRE_STREAM_UNKNOWN : e;
}

void sre_stream_fn_compose(uint8_t *fn, const uint8_t *then)
{
    rec_t a, b;
    memcpy(a.w, fn, sizeof(a.w));
    memcpy(b.w, then, sizeof(b.w));
    a.then(b);
    memcpy(fn, a.w, sizeof(a.w));
}

void sre_stream_fn_identity(uint8_t *fn)
{
    rec_t a;
    a.identity();
    memcpy(fn, a.w, sizeof(a.w));
}

int sre_stream_fn_unresolved(const uint8_t *fn)
{
    rec_t r;
    memcpy(r.w, fn, sizeof(r.w));
    return r.is_unresolved() ? 1 : 0;
}

/*
 * Reduce the stream part to its records (all levels).  ws.count[] / ws.fn[]
 * are sized by the caller: count[0] = ceil(len / PIECE) (1 when len == 0),
 * count[l+1] = ceil(count[l] / FAN) while count[l] > FAN.
 */
cudaError_t sre_launch_dfa_stream_reduce(const sre_dev_dfa_t &dfa, const sre_dev_image_t &img,
    const uint8_t *buf, size_t len, const uint8_t *halo, uint32_t entry, const sre_stream_ws_t &ws,
    cudaStream_t stream, int *launches)
{
    cudaError_t err;
    const size_t nfull = len / PIECE, tail = len % PIECE;
    const stream_src_t src = { buf, len, halo, entry };
    if (dfa.acc != ACC || img.K != (uint32_t) K) {
        return cudaErrorInvalidValue;
    }
    if (nfull) {
        if (launches) ++*launches;
        /* [state][byte] rows in shared memory for byte-table DFAs; the class table in
         * shared memory while it fits next to the staging rings, else through L1/L2 */
        const size_t room = SMEM_LIMIT - 32 * 32 * 128 - 4096;
        if (dfa.t256 != nullptr) {
            err = launch_pieces<0>(dfa, img, src, nfull, ws.fn[0], stream);
        } else if ((size_t) dfa.nstates * dfa.nclasses * 2 <= room) {
            err = launch_pieces<1>(dfa, img, src, nfull, ws.fn[0], stream);
        } else {
            err = launch_pieces<2>(dfa, img, src, nfull, ws.fn[0], stream);
        }
        if (err != cudaSuccess) return err;
    }
    if (tail || len == 0) {
        if (launches) ++*launches;
        k_stream_tail<<<1, 32, 0, stream>>>(dfa, img, src, nfull * PIECE, ws.fn[0] + nfull * sizeof(rec_t));
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    for (int l = 0; l < 3 && ws.count[l] > FAN; l++) {
        const size_t n_out = ws.count[l + 1];
        if (launches) ++*launches;
        k_stream_compose<<<(unsigned) ((n_out * 32 + 255) / 256), 256, 0, stream>>>(
            ws.fn[l], ws.count[l], ws.fn[l + 1], n_out, nullptr, 0);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

/*
 * One round of repair: resolve the first unresolved piece (if any) with the
 * stream's true entry state and recompose its ancestors.  *dev_fixed = that
 * piece, or ~0 when all records were resolved already.
 */
cudaError_t sre_launch_dfa_stream_fix(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
    uint32_t entry_state, const sre_stream_ws_t &ws, unsigned long long *dev_fixed, cudaStream_t stream,
    int *launches)
{
    const stream_levels_t lv = levels_of(ws);
    const stream_src_t src = { buf, len, nullptr, entry_state };
    if (launches) ++*launches;
    k_stream_fix<<<1, 32, 0, stream>>>(dfa, src, lv, entry_state, dev_fixed);
    cudaError_t err = cudaGetLastError();
    unsigned long long span = FAN;
    for (int l = 0; l < lv.top && err == cudaSuccess; l++, span *= FAN) {
        if (launches) ++*launches;
        k_stream_compose<<<1, 32, 0, stream>>>(ws.fn[l], ws.count[l], ws.fn[l + 1], ws.count[l + 1], dev_fixed, span);
        err = cudaGetLastError();
    }
    return err;
}

/*
 * Walk the (fully resolved) levels from the true entry state: dev_out[0] = state
 * after the stream part (0xffffffff: a record did not know the state -- the
 * caller treats that as an internal error), *dev_match_offset = offset of the
 * step that enters ACC, or -1.
 */
cudaError_t sre_launch_dfa_stream_walk(const sre_dev_dfa_t &dfa, const sre_dev_image_t &img,
    const uint8_t *buf, size_t len, uint32_t entry_state, const sre_stream_ws_t &ws, uint32_t *dev_out,
    long long *dev_match_offset, cudaStream_t stream, int *launches)
{
    const stream_levels_t lv = levels_of(ws);
    const stream_src_t src = { buf, len, nullptr, SRE_STREAM_UNKNOWN };
    if (launches) *launches += 2;
    k_stream_descend<<<1, 32, 0, stream>>>(lv, entry_state, ws.first_acc, dev_out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        return err;
    }
    k_stream_locate<<<1, 32, 0, stream>>>(dfa, img, src, ws.first_acc, dev_out, dev_match_offset);
    return cudaGetLastError();
}

/* restart point of a Pike search that must find the match the scan saw at step `limit - 1`
 * (or at the EOF step: limit = len); the records must be resolved (sre_launch_dfa_stream_walk ran) */
cudaError_t sre_launch_dfa_stream_restart(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
    uint32_t entry_state, size_t limit, const sre_stream_ws_t &ws, long long *dev_out, cudaStream_t stream,
    int *launches)
{
    const stream_levels_t lv = levels_of(ws);
    const stream_src_t src = { buf, len, nullptr, entry_state };
    if (launches) ++*launches;
    k_stream_restart<<<1, 32, 0, stream>>>(dfa, src, lv, entry_state, limit, 256, dev_out);
    return cudaGetLastError();
}
