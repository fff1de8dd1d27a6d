/*
 * sre_stream.cu -- chunk-parallel Thompson matching of ONE long stream.
 *
 * What it replaces: a sequence of sre_vm_thompson_exec(ctx, chunk, n, eof)
 * calls with SRE_AGAIN state carry (reference sre_vm_thompson.c:63-270, ctx
 * fields sre_vm_thompson.h:30-41).  The carried state of the determinised
 * program is one DFA state, so a piece of the stream is a function
 * f: state -> state, and function composition is associative:
 *
 *   k_stream_pieces   one CUDA thread per PIECE-byte piece (same TMA tile
 *                     pipeline as k_dfa_lines, rows = pieces): runs the first
 *                     16-byte chunks of the piece from EVERY entry state
 *                     ("speculatively") until all of them have either been
 *                     absorbed by ACC or met in one state, then the single-state
 *                     inner loop of k_dfa_lines; writes the piece's transfer
 *                     function (nstates bytes, padded to 16/32).
 *   k_stream_tail     the ragged end (< PIECE bytes): one warp, 128-byte
 *                     sub-chunks per lane, shuffle-tree composition.
 *   k_stream_compose  one warp composes FAN = 256 consecutive functions (8 per
 *                     lane serially, then a 5-round shuffle tree) into one
 *                     function of the next level; loads are coalesced.
 *   k_stream_descend  one warp walks back down from the true entry state along
 *                     the only path that matters: at each level the child of
 *                     the current parent whose step first enters ACC (= in
 *                     which the reference's loop would return SRE_OK); also
 *                     yields the state after the whole stream.
 *   k_stream_locate   one warp re-runs that piece to get the exact byte offset.
 *
 * Everything is exact: no result depends on a guess.
 */
#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

uint32_t g_piece = 4096;            /* bytes per level-0 piece (1024..8192, power of two) */
constexpr uint32_t FAN = 256;       /* functions composed per warp            */
constexpr uint32_t PER_LANE = FAN / 32;

/* function records are 16 or 32 bytes (nstates rounded up) */
__host__ __device__ inline uint32_t fn_stride(uint32_t nstates)
{
    return nstates <= 16 ? 16 : 32;
}

/* ---- a state -> state function packed into NW = DV/4 words --------------------- */

template <int NW>
struct fn_t {
    uint32_t w[NW];

    __device__ __forceinline__ void identity()
    {
#pragma unroll
        for (int i = 0; i < NW; i++) {
            w[i] = 0x03020100u + 0x04040404u * i;
        }
    }
    /* byte k, k dynamic */
    __device__ __forceinline__ uint32_t at(uint32_t k) const
    {
        uint32_t v = w[0];
#pragma unroll
        for (int i = 1; i < NW; i++) {
            v = (k >> 2) == (uint32_t) i ? w[i] : v;
        }
        return (v >> ((k & 3) * 8)) & 0xff;
    }
    /* byte d, d a compile-time constant after unrolling */
    __device__ __forceinline__ uint32_t get(int d) const { return (w[d >> 2] >> ((d & 3) * 8)) & 0xff; }
    __device__ __forceinline__ void set(int d, uint32_t v)
    {
        w[d >> 2] = (w[d >> 2] & ~(0xffu << ((d & 3) * 8))) | (v << ((d & 3) * 8));
    }
    /* this = g o this  (apply this first, then g) */
    __device__ __forceinline__ void then(const fn_t &g)
    {
        fn_t r;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            r.w[i] = 0;
        }
#pragma unroll
        for (int d = 0; d < NW * 4; d++) {
            r.w[d >> 2] |= g.at(get(d)) << ((d & 3) * 8);
        }
        *this = r;
    }
    __device__ __forceinline__ void load(const uint8_t *p)
    {
        const uint4 a = *reinterpret_cast<const uint4 *>(p);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        if (NW > 4) {
            const uint4 b = *reinterpret_cast<const uint4 *>(p + 16);
            w[4 % NW] = b.x; w[5 % NW] = b.y; w[6 % NW] = b.z; w[7 % NW] = b.w;
        }
    }
    __device__ __forceinline__ void store(uint8_t *p) const
    {
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
        if (NW > 4) {
            *reinterpret_cast<uint4 *>(p + 16) = make_uint4(w[4 % NW], w[5 % NW], w[6 % NW], w[7 % NW]);
        }
    }
    __device__ __forceinline__ fn_t shfl_down(uint32_t delta) const
    {
        fn_t r;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            r.w[i] = __shfl_down_sync(0xffffffffu, w[i], delta);
        }
        return r;
    }
};

/* compose the functions held by the 32 lanes in lane order; lane 0 gets the result */
template <int NW>
__device__ __forceinline__ void warp_compose(fn_t<NW> &f)
{
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (uint32_t s = 1; s < 32; s <<= 1) {
        const fn_t<NW> g = f.shfl_down(s);
        if ((lane & (2 * s - 1)) == 0) {
            f.then(g);
        }
    }
}

/* ---- level 0: pieces ------------------------------------------------------------ */

/* NW*4-way consumer: byte d of `st` = state reached from entry state d */
template <int NW>
struct piece_consumer_t {
    uint32_t        tab_s;      /* [nstates] rows of ROW260 bytes in shared memory (shared-window address) */
    uint8_t        *fn;         /* level-0 function records                   */
    uint32_t        nstates, acc;
    size_t          npieces;
    fn_t<NW>        st;
    uint32_t        s;          /* the single live state once converged       */
    uint32_t        accmask;    /* entry states already absorbed by ACC       */
    bool            conv;

    __device__ __forceinline__ void begin()
    {
        st.identity();
        conv = false;
        s = 0;
        accmask = 0;
    }

    /* converged = every entry state was absorbed by ACC (e.g. entry states that
     * already hold a MATCH thread) or has reached one common state */
    __device__ __forceinline__ void check()
    {
        uint32_t ref = acc, mask = 0;
        bool all = true;
#pragma unroll
        for (int d = 0; d < NW * 4; d++) {
            if (d < (int) nstates) {
                const uint32_t v = st.get(d);
                if (v == acc) {
                    mask |= 1u << d;
                } else if (ref == acc) {
                    ref = v;
                } else if (v != ref) {
                    all = false;
                }
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
        step260_t st256 = { tab_s };
        if (conv) {
            s = st256.word(st256.word(st256.word(st256.word(s, v.x), v.y), v.z), v.w);
            return;
        }
#pragma unroll
        for (int d = 0; d < NW * 4; d++) {
            if (d < (int) nstates) {
                uint32_t x = st.get(d);
                x = st256.word(st256.word(st256.word(st256.word(x, v.x), v.y), v.z), v.w);
                st.set(d, x);
            }
        }
        check();
    }

    __device__ __forceinline__ void byte(uint32_t b)
    {
        step260_t st256 = { tab_s };
        if (conv) {
            s = st256.byte(s, b);
            return;
        }
#pragma unroll
        for (int d = 0; d < NW * 4; d++) {
            if (d < (int) nstates) {
                st.set(d, st256.byte(st.get(d), b));
            }
        }
    }

    __device__ __forceinline__ void end(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        if (piece >= npieces) {
            return;
        }
        if (conv) {
#pragma unroll
            for (int d = 0; d < NW * 4; d++) {
                st.set(d, d < (int) nstates ? (((accmask >> d) & 1) ? acc : s) : 0);
            }
        } else {
#pragma unroll
            for (int d = 0; d < NW * 4; d++) {
                if (d >= (int) nstates) {
                    st.set(d, 0);
                }
            }
        }
        st.store(fn + piece * (NW * 4));
    }
};

template <int NW>
__global__ void __launch_bounds__(1024, 1)
k_stream_pieces(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, size_t npieces, uint8_t *fn,
                uint32_t PIECE)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    /* one extra row of room for the padding of the rows */
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates + 1, dfa.nclasses, false);
    load_table260(smem, dfa.t256, dfa.nstates);
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    piece_consumer_t<NW> cons;
    cons.tab_s = (uint32_t) __cvta_generic_to_shared(smem);
    cons.fn = fn;
    cons.nstates = dfa.nstates;
    cons.acc = dfa.acc;
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

/* function of bytes [begin, end) computed from every entry state (global table) */
template <int NW>
__device__ __forceinline__ void span_function(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t begin,
                                              size_t end, fn_t<NW> &f)
{
    f.identity();
    for (size_t i = begin; i < end; i++) {
        const uint32_t b = buf[i];
#pragma unroll
        for (int d = 0; d < NW * 4; d++) {
            if (d < (int) dfa.nstates) {
                f.set(d, __ldg(dfa.t256 + ((f.get(d) << 8) | b)));
            }
        }
    }
#pragma unroll
    for (int d = 0; d < NW * 4; d++) {
        if (d >= (int) dfa.nstates) {
            f.set(d, 0);
        }
    }
}

/* the ragged tail (< PIECE bytes): one warp, 128-byte sub-chunks per lane */
template <int NW>
__global__ void __launch_bounds__(32)
k_stream_tail(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len, uint8_t *out)
{
    const uint32_t lane = threadIdx.x;
    const size_t sub = (len + 31) / 32;
    const size_t b = lane * sub < len ? lane * sub : len, e = b + sub < len ? b + sub : len;
    fn_t<NW> f;
    span_function<NW>(dfa, buf, b, e, f);
    warp_compose<NW>(f);
    if (lane == 0) {
        f.store(out);
    }
}

/* ---- upper levels ------------------------------------------------------------------ */

/* lane's local composition of its PER_LANE consecutive records */
template <int NW>
__device__ __forceinline__ void lane_compose(const uint8_t *in, size_t first, size_t n_in, fn_t<NW> &f)
{
    f.identity();
#pragma unroll
    for (uint32_t k = 0; k < PER_LANE; k++) {
        if (first + k < n_in) {
            fn_t<NW> g;
            g.load(in + (first + k) * (NW * 4));
            f.then(g);
        }
    }
}

/* out[w] = in[w*FAN + FAN-1] o ... o in[w*FAN], one warp per output */
template <int NW>
__global__ void __launch_bounds__(256)
k_stream_compose(const uint8_t *__restrict__ in, size_t n_in, uint8_t *__restrict__ out, size_t n_out)
{
    const size_t w = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n_out) {
        return;
    }
    fn_t<NW> f;
    lane_compose<NW>(in, w * FAN + lane * PER_LANE, n_in, f);
    warp_compose<NW>(f);
    if (lane == 0) {
        f.store(out + w * (NW * 4));
    }
}

/*
 * The walk down needs one path only: at every level, the child of the current
 * parent whose step takes a non-ACC state into ACC (ACC is absorbing, so that
 * is the child in which the match is first seen).  One warp: per level it
 * composes the <= FAN children of the parent (8 per lane), propagates the entry
 * state across the lanes and picks the first such child.  Outputs: the state
 * after the whole stream, the first piece whose step enters ACC (or ~0) and the
 * exact entry state of that piece (entry0[piece], for k_stream_locate).
 */
struct stream_levels_t {
    const uint8_t *fn[4];
    size_t         count[4];
    int            top;
};

template <int NW>
__global__ void __launch_bounds__(32)
k_stream_descend(stream_levels_t lv, uint32_t root_state, uint32_t acc, uint8_t *entry0,
                 unsigned long long *first_acc, uint32_t *exit_state)
{
    const uint32_t lane = threadIdx.x;
    size_t parent = 0;
    uint32_t s_in = root_state;
    for (int l = lv.top; l >= 0; l--) {
        const size_t n = lv.count[l], first = parent * FAN + lane * PER_LANE;
        fn_t<NW> total;
        lane_compose<NW>(lv.fn[l], first, n, total);
        uint32_t s = s_in, mine = 0;
#pragma unroll 4
        for (uint32_t k = 0; k < 32; k++) {
            if (lane == k) {
                mine = s;
            }
            s = __shfl_sync(0xffffffffu, total.at(s), k);
        }
        if (l == lv.top && lane == 0 && exit_state) {
            *exit_state = s;
        }
        s = mine;
        unsigned long long hit = ~0ull;
        uint32_t hit_entry = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER_LANE; k++) {
            const size_t j = first + k;
            if (j < n && j < (parent + 1) * FAN) {
                fn_t<NW> g;
                g.load(lv.fn[l] + j * (NW * 4));
                const uint32_t nx = g.at(s);
                if (nx == acc && s != acc && hit == ~0ull) {
                    hit = j;
                    hit_entry = s;
                }
                s = nx;
            }
        }
        const uint32_t who = __ffs(__ballot_sync(0xffffffffu, hit != ~0ull));
        if (who == 0) {
            if (lane == 0) {
                *first_acc = ~0ull;
            }
            return;
        }
        parent = (size_t) __shfl_sync(0xffffffffu, hit, who - 1);
        s_in = __shfl_sync(0xffffffffu, hit_entry, who - 1);
    }
    if (lane == 0) {
        *first_acc = (unsigned long long) parent;
        entry0[parent] = (uint8_t) s_in;
    }
}

/* exact offset of the byte whose step enters ACC inside piece *first_acc */
template <int NW>
__global__ void __launch_bounds__(32)
k_stream_locate(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len,
                const uint8_t *__restrict__ entry0, const unsigned long long *first_acc,
                long long *match_offset, uint32_t PIECE)
{
    const uint32_t lane = threadIdx.x;
    const unsigned long long piece = *first_acc;
    if (piece == ~0ull) {
        if (lane == 0) {
            *match_offset = -1;
        }
        return;
    }
    const size_t start = (size_t) piece * PIECE, end = start + PIECE < len ? start + PIECE : len;
    const size_t sub = (end - start + 31) / 32;
    const size_t b = start + lane * sub < end ? start + lane * sub : end, e = b + sub < end ? b + sub : end;
    fn_t<NW> f;
    span_function<NW>(dfa, buf, b, e, f);

    uint32_t s = entry0[piece], mine = 0;
    for (uint32_t l = 0; l < 32; l++) {
        if (lane == l) {
            mine = s;
        }
        s = __shfl_sync(0xffffffffu, f.at(s), l);
    }
    /* the first lane whose sub-chunk takes a non-ACC state into ACC */
    const bool hit = mine != dfa.acc && f.at(mine) == dfa.acc;
    const uint32_t who = __ffs(__ballot_sync(0xffffffffu, hit));
    if (who == 0) {
        if (lane == 0) {
            *match_offset = -1;     /* cannot happen */
        }
        return;
    }
    if (lane == who - 1) {
        s = mine;
        for (size_t i = b; i < e; i++) {
            s = __ldg(dfa.t256 + ((s << 8) | buf[i]));
            if (s == dfa.acc) {
                *match_offset = (long long) i;
                return;
            }
        }
        *match_offset = -1;
    }
}

template <int NW>
cudaError_t launch_pieces(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t npieces, uint8_t *fn,
    cudaStream_t stream)
{
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates + 1, dfa.nclasses, false);   /* as the kernel */
    const int warps = 32;
    const size_t smem = plan.stage_ofs + (size_t) warps * 32 * 128;
    const uint32_t PIECE = g_piece;
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, buf, npieces, PIECE, 128);
    if (err != cudaSuccess) {
        return err;
    }
    auto kern = k_stream_pieces<NW>;
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
    kern<<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, tmap, npieces, fn, PIECE);
    return cudaGetLastError();
}

int top_level(const sre_stream_ws_t &ws)
{
    int top = 0;
    while (top < 3 && ws.count[top] > FAN) {
        top++;
    }
    return top;
}

template <int NW>
cudaError_t reduce_t(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len, const sre_stream_ws_t &ws,
    cudaStream_t stream, int *launches)
{
    cudaError_t err;
    const uint32_t fs = NW * 4, PIECE = g_piece;
    const size_t nfull = len / PIECE, tail = len % PIECE;
    if (nfull) {
        if (launches) ++*launches;
        if ((err = launch_pieces<NW>(dfa, buf, nfull, ws.fn[0], stream)) != cudaSuccess) return err;
    }
    if (tail || len == 0) {
        if (launches) ++*launches;
        k_stream_tail<NW><<<1, 32, 0, stream>>>(dfa, buf + nfull * PIECE, tail, ws.fn[0] + nfull * fs);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    for (int l = 0; l < 3 && ws.count[l] > FAN; l++) {
        const size_t n_out = ws.count[l + 1];
        if (launches) ++*launches;
        k_stream_compose<NW><<<(unsigned) ((n_out * 32 + 255) / 256), 256, 0, stream>>>(
            ws.fn[l], ws.count[l], ws.fn[l + 1], n_out);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

template <int NW>
cudaError_t walk_t(const sre_dev_dfa_t &dfa, uint32_t entry_state, const sre_stream_ws_t &ws,
    uint32_t *exit_state, cudaStream_t stream, int *launches)
{
    stream_levels_t lv;
    lv.top = top_level(ws);
    for (int l = 0; l < 4; l++) {
        lv.fn[l] = ws.fn[l];
        lv.count[l] = ws.count[l];
    }
    if (launches) ++*launches;
    k_stream_descend<NW><<<1, 32, 0, stream>>>(lv, entry_state, dfa.acc, ws.entry[0], ws.first_acc, exit_state);
    return cudaGetLastError();
}

}  // namespace

size_t sre_stream_piece_bytes(void) { return g_piece; }
void sre_stream_set_piece_bytes(uint32_t b)
{
    if (b == 1024 || b == 2048 || b == 4096 || b == 8192) {
        g_piece = b;
    }
}
uint32_t sre_stream_fan(void) { return FAN; }
uint32_t sre_stream_fn_stride(uint32_t nstates) { return fn_stride(nstates); }

/*
 * Reduce the stream to its transfer functions (all levels).  ws.count[] /
 * ws.fn[] / ws.entry[] are sized by the caller: count[0] = ceil(len / PIECE)
 * (1 when len == 0), count[l+1] = ceil(count[l] / FAN) while count[l] > FAN.
 */
cudaError_t sre_launch_dfa_stream_reduce(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
    const sre_stream_ws_t &ws, cudaStream_t stream, int *launches)
{
    return dfa.nstates <= 16 ? reduce_t<4>(dfa, buf, len, ws, stream, launches)
                             : reduce_t<8>(dfa, buf, len, ws, stream, launches);
}

/*
 * Walk the reduced levels from the true entry state: exact entry state of every
 * piece, *exit_state = state after the stream, ws.first_acc = first piece whose
 * step enters ACC (or ~0).
 */
cudaError_t sre_launch_dfa_stream_walk(const sre_dev_dfa_t &dfa, uint32_t entry_state,
    const sre_stream_ws_t &ws, uint32_t *exit_state, cudaStream_t stream, int *launches)
{
    return dfa.nstates <= 16 ? walk_t<4>(dfa, entry_state, ws, exit_state, stream, launches)
                             : walk_t<8>(dfa, entry_state, ws, exit_state, stream, launches);
}

cudaError_t sre_launch_dfa_stream_locate(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
    const sre_stream_ws_t &ws, long long *dev_match_offset, cudaStream_t stream, int *launches)
{
    if (launches) ++*launches;
    if (dfa.nstates <= 16) {
        k_stream_locate<4><<<1, 32, 0, stream>>>(dfa, buf, len, ws.entry[0], ws.first_acc, dev_match_offset,
                                                 g_piece);
    } else {
        k_stream_locate<8><<<1, 32, 0, stream>>>(dfa, buf, len, ws.entry[0], ws.first_acc, dev_match_offset,
                                                 g_piece);
    }
    return cudaGetLastError();
}
