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
 *                     pipeline as k_dfa_lines, rows = pieces): runs the
 *                     piece from EVERY entry state at once ("speculatively")
 *                     and writes the piece's transfer function (nstates bytes).
 *                     As soon as all live entry states have converged to one
 *                     state the thread drops to the single-state inner loop, so
 *                     for automata that forget their history quickly the cost
 *                     is ~1 table look-up per byte, as in k_dfa_lines.
 *   k_stream_compose  thread i composes G consecutive functions of one level
 *                     into one function of the next level (reduction tree).
 *   k_stream_top      one thread walks the (<= G) functions of the top level
 *                     from the true entry state: exact entry state of every
 *                     top-level element, and the exit state of the stream.
 *   k_stream_entries  pushes exact entry states down one level; at level 0 it
 *                     records the first piece in which the automaton enters
 *                     ACC (= in which the reference's loop would return SRE_OK).
 *   k_stream_locate   re-runs that one piece to get the exact byte offset.
 *
 * Everything is exact: no piece result depends on a guess.
 */
#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

constexpr uint32_t PIECE = 1024;    /* bytes per level-0 piece                */
constexpr uint32_t FAN = 256;       /* functions composed per thread          */

/* function records are FS = 16/32 bytes apart (nstates rounded up) */
__host__ __device__ inline uint32_t fn_stride(uint32_t nstates)
{
    return nstates <= 16 ? 16 : 32;
}

/* D-way consumer: st[d] = state reached from entry state d */
template <int DV>
struct piece_consumer_t {
    const uint8_t  *tab;        /* [nstates][256] in shared memory            */
    uint8_t        *fn;         /* level-0 function records                   */
    uint32_t        nstates, acc, fs;
    size_t          npieces;
    uint32_t        st[DV];
    uint32_t        s;          /* the single state once converged            */
    uint32_t        accmask;    /* entry states already absorbed by ACC       */
    bool            conv;

    __device__ __forceinline__ void begin()
    {
#pragma unroll
        for (int d = 0; d < DV; d++) {
            st[d] = d < (int) nstates ? d : 0;
        }
        conv = false;
        s = 0;
        accmask = 0;
    }

    __device__ __forceinline__ void check()
    {
        /* converged = every entry state has either been absorbed by ACC (e.g.
         * entry states that already hold a MATCH thread) or reached one common
         * state */
        uint32_t ref = acc, mask = 0;
        bool all = true;
#pragma unroll
        for (int d = 0; d < DV; d++) {
            if (d < (int) nstates) {
                if (st[d] == acc) {
                    mask |= 1u << d;
                } else if (ref == acc) {
                    ref = st[d];
                } else if (st[d] != ref) {
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
        if (conv) {
            step256_t st256 = { tab };
            s = st256.word(st256.word(st256.word(st256.word(s, v.x), v.y), v.z), v.w);
            return;
        }
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t b = (w[i] >> (8 * k)) & 0xff;
#pragma unroll
                for (int d = 0; d < DV; d++) {
                    st[d] = tab[(st[d] << 8) | b];
                }
            }
        }
        check();
    }

    __device__ __forceinline__ void byte(uint32_t b)
    {
        if (conv) {
            s = tab[(s << 8) | b];
            return;
        }
#pragma unroll
        for (int d = 0; d < DV; d++) {
            st[d] = tab[(st[d] << 8) | b];
        }
    }

    __device__ __forceinline__ void end(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        if (piece >= npieces) {
            return;
        }
        uint8_t *out = fn + piece * fs;
#pragma unroll
        for (int d = 0; d < DV; d++) {
            if (d < (int) fs) {
                const uint32_t v = conv ? (((accmask >> d) & 1) ? acc : s) : st[d];
                out[d] = (uint8_t) (d < (int) nstates ? v : 0);
            }
        }
    }
};

template <int DV, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
k_stream_pieces(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, size_t npieces, uint8_t *fn)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, false);
    load_table(smem, dfa.t256, plan.tab_bytes);
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    piece_consumer_t<DV> cons;
    cons.tab = smem;
    cons.fn = fn;
    cons.nstates = dfa.nstates;
    cons.acc = dfa.acc;
    cons.fs = fn_stride(dfa.nstates);
    cons.npieces = npieces;

    /* rows = pieces: the stream is a {PIECE, npieces} byte tensor */
    tile_pipeline_tma_early<1>(cons, &tmap, npieces, PIECE,
                               smem + plan.stage_ofs + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + plan.bar_ofs) + warp * MAX_STAGES,
                               (size_t) blockIdx.x * warps_per_block + warp,
                               (size_t) gridDim.x * warps_per_block);
}

/* transfer function of the ragged tail (< PIECE bytes), one thread per state */
__global__ void k_stream_tail(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len, uint8_t *out)
{
    const uint32_t d = threadIdx.x;
    if (d >= fn_stride(dfa.nstates)) {
        return;
    }
    uint32_t s = d < dfa.nstates ? d : 0;
    for (size_t i = 0; i < len; i++) {
        s = dfa.t256[(s << 8) | buf[i]];
    }
    out[d] = (uint8_t) (d < dfa.nstates ? s : 0);
}

/* a function record held in registers: byte k of a 16/32-byte record */
struct fn_rec_t {
    uint4 lo, hi;
    __device__ __forceinline__ void load(const uint8_t *p, uint32_t fs)
    {
        lo = *reinterpret_cast<const uint4 *>(p);
        if (fs > 16) {
            hi = *reinterpret_cast<const uint4 *>(p + 16);
        }
    }
    __device__ __forceinline__ uint32_t at(uint32_t k) const
    {
        const uint4 &q = (k & 16) ? hi : lo;
        const uint32_t w = (k & 8) ? ((k & 4) ? q.w : q.z) : ((k & 4) ? q.y : q.x);
        return (w >> ((k & 3) * 8)) & 0xff;
    }
};

/* out[i] = in[i*FAN + last] o ... o in[i*FAN].  The record loads do not depend
 * on the running composition, so they pipeline; only register selects do. */
template <int DV>
__global__ void __launch_bounds__(128)
k_stream_compose(const uint8_t *__restrict__ in, size_t n_in, uint8_t *__restrict__ out, size_t n_out,
                 uint32_t nstates, uint32_t fs)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) {
        return;
    }
    uint32_t cur[DV];
#pragma unroll
    for (int d = 0; d < DV; d++) {
        cur[d] = d;
    }
    const size_t first = i * FAN, last = first + FAN < n_in ? first + FAN : n_in;
#pragma unroll 4
    for (size_t j = first; j < last; j++) {
        fn_rec_t f;
        f.load(in + j * fs, fs);
#pragma unroll
        for (int d = 0; d < DV; d++) {
            cur[d] = f.at(cur[d]);
        }
    }
    uint8_t *o = out + i * fs;
#pragma unroll
    for (int d = 0; d < DV; d++) {
        o[d] = d < (int) nstates ? (uint8_t) cur[d] : 0;
    }
    for (uint32_t d = DV; d < fs; d++) {
        o[d] = 0;
    }
}

/* serial walk of the top level: entry state of each element, stream exit */
__global__ void k_stream_top(const uint8_t *__restrict__ fn, size_t n, uint32_t fs, uint32_t entry_state,
                             uint8_t *entry, uint32_t *exit_state, unsigned long long *first_acc)
{
    uint32_t s = entry_state;
    for (size_t j = 0; j < n; j++) {
        entry[j] = (uint8_t) s;
        s = fn[j * fs + s];
    }
    *exit_state = s;
    *first_acc = ~0ull;
}

/* entries of level L from entries of level L+1; level 0 records first ACC */
__global__ void __launch_bounds__(128)
k_stream_entries(const uint8_t *__restrict__ fn, size_t n, uint32_t fs, const uint8_t *__restrict__ parent_entry,
                 size_t n_parent, uint8_t *entry, uint32_t acc, int level0, unsigned long long *first_acc)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parent) {
        return;
    }
    uint32_t s = parent_entry[i];
    const size_t first = i * FAN, last = first + FAN < n ? first + FAN : n;
#pragma unroll 4
    for (size_t j = first; j < last; j++) {
        fn_rec_t f;
        f.load(fn + j * fs, fs);
        if (entry) {
            entry[j] = (uint8_t) s;
        }
        const uint32_t nx = f.at(s);
        if (level0 && nx == acc && s != acc) {
            atomicMin(first_acc, (unsigned long long) j);
        }
        s = nx;
    }
}

/* exact offset (relative to buf) of the byte whose step enters ACC */
__global__ void k_stream_locate(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len,
                                const uint8_t *__restrict__ entry0, uint32_t entry_state_if_single,
                                const unsigned long long *first_acc, long long *match_offset)
{
    const unsigned long long piece = *first_acc;
    if (piece == ~0ull) {
        *match_offset = -1;
        return;
    }
    uint32_t s = entry0 ? entry0[piece] : entry_state_if_single;
    const size_t start = (size_t) piece * PIECE, end = start + PIECE < len ? start + PIECE : len;
    for (size_t i = start; i < end; i++) {
        s = dfa.t256[(s << 8) | buf[i]];
        if (s == dfa.acc) {
            *match_offset = (long long) i;
            return;
        }
    }
    *match_offset = -1;     /* cannot happen */
}

template <int DV, int THREADS>
cudaError_t launch_pieces(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t npieces, uint8_t *fn,
    cudaStream_t stream)
{
    const dfa_smem_plan_t plan = dfa_smem_plan(dfa.nstates, dfa.nclasses, false);
    const int warps = THREADS / 32;
    const size_t smem = plan.stage_ofs + (size_t) warps * 32 * 128;
    CUtensorMap tmap;
    cudaError_t err = make_row_tensor_map(&tmap, buf, npieces, PIECE, 128);
    if (err != cudaSuccess) {
        return err;
    }
    auto kern = k_stream_pieces<DV, THREADS>;
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
    kern<<<(unsigned) grid, THREADS, smem, stream>>>(dfa, tmap, npieces, fn);
    return cudaGetLastError();
}

}  // namespace

size_t sre_stream_piece_bytes(void) { return PIECE; }
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
    cudaError_t err;
    const uint32_t fs = fn_stride(dfa.nstates);
    const size_t nfull = len / PIECE, tail = len % PIECE;

    if (nfull) {
        if (launches) ++*launches;
        if (dfa.nstates <= 8) {
            err = launch_pieces<8, 1024>(dfa, buf, nfull, ws.fn[0], stream);
        } else if (dfa.nstates <= 16) {
            err = launch_pieces<16, 768>(dfa, buf, nfull, ws.fn[0], stream);
        } else {
            err = launch_pieces<32, 512>(dfa, buf, nfull, ws.fn[0], stream);
        }
        if (err != cudaSuccess) return err;
    }
    if (tail || len == 0) {
        if (launches) ++*launches;
        k_stream_tail<<<1, 32, 0, stream>>>(dfa, buf + nfull * PIECE, tail, ws.fn[0] + nfull * fs);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    for (int l = 0; l < 3 && ws.count[l] > FAN; l++) {
        const size_t n_out = ws.count[l + 1];
        if (launches) ++*launches;
        if (dfa.nstates <= 16) {
            k_stream_compose<16><<<(unsigned) ((n_out + 127) / 128), 128, 0, stream>>>(
                ws.fn[l], ws.count[l], ws.fn[l + 1], n_out, dfa.nstates, fs);
        } else {
            k_stream_compose<32><<<(unsigned) ((n_out + 127) / 128), 128, 0, stream>>>(
                ws.fn[l], ws.count[l], ws.fn[l + 1], n_out, dfa.nstates, fs);
        }
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

/*
 * Walk the reduced levels from the true entry state: exact entry state of every
 * piece, *exit_state = state after the stream, ws.first_acc = first piece whose
 * step enters ACC (or ~0).
 */
cudaError_t sre_launch_dfa_stream_walk(const sre_dev_dfa_t &dfa, uint32_t entry_state,
    const sre_stream_ws_t &ws, uint32_t *exit_state, cudaStream_t stream, int *launches)
{
    cudaError_t err;
    const uint32_t fs = fn_stride(dfa.nstates);
    int top = 0;
    while (top < 3 && ws.count[top] > FAN) {
        top++;
    }
    if (launches) ++*launches;
    k_stream_top<<<1, 1, 0, stream>>>(ws.fn[top], ws.count[top], fs, entry_state, ws.entry[top],
                                      exit_state, ws.first_acc);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;

    if (top == 0) {
        /* single level (count[0] <= FAN): one thread looks for the first ACC */
        if (launches) ++*launches;
        k_stream_entries<<<1, 1, 0, stream>>>(ws.fn[0], ws.count[0], fs, ws.entry[0], 1, nullptr, dfa.acc,
                                              1, ws.first_acc);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    for (int l = top - 1; l >= 0; l--) {
        const size_t n_parent = ws.count[l + 1];
        if (launches) ++*launches;
        k_stream_entries<<<(unsigned) ((n_parent + 127) / 128), 128, 0, stream>>>(
            ws.fn[l], ws.count[l], fs, ws.entry[l + 1], n_parent, ws.entry[l], dfa.acc, l == 0,
            ws.first_acc);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

cudaError_t sre_launch_dfa_stream_locate(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
    const sre_stream_ws_t &ws, long long *dev_match_offset, cudaStream_t stream, int *launches)
{
    if (launches) ++*launches;
    k_stream_locate<<<1, 1, 0, stream>>>(dfa, buf, len, ws.entry[0], 0, ws.first_acc, dev_match_offset);
    return cudaGetLastError();
}
