/*
 * sre_text.cu -- Thompson matching of every line of a '\n'-delimited buffer in
 * ONE pass over the text (what grep does with a log file).
 *
 * What it replaces: sre_vm_thompson_exec(ctx, line, len, eof=1) with a fresh
 * context per line (sre_vm_thompson.c:63-270), for lines that are not laid out
 * at a fixed pitch.  The older route -- sre_cuda_index_lines (two passes) and
 * then the ragged thread-per-line kernel (uncoalesced) -- read the text three
 * times at ~0.7 TB/s.  Here the text goes once through the TMA tile pipeline
 * of k_dfa_lines_tma_early, cut into PIECE-byte pieces (one CUDA thread each):
 *
 * With line offsets wanted (rc[line] and offsets[line + 1]):
 *
 *   k_text_pieces   The automaton table has the line structure folded in: on
 *                   '\n' every state goes to the START state, with bit 7 set
 *                   when the line that just ended matched (the verdict of the
 *                   reference's EOF step after the terminator; rows r and r+128
 *                   of the table are the same row).  So a thread simply runs its
 *                   piece from the start state: whatever it computes before the
 *                   first '\n' is thrown away (that line began in an earlier
 *                   piece and belongs to the thread of that piece), every later
 *                   line is exact, and the line that is still open at the end
 *                   of the piece is finished by reading on past it.  Line ends
 *                   and verdicts go to a small per-piece staging area.
 *   k_text_sums / _scan   exclusive scan of the per-piece line counts.
 *   k_text_write    staging -> rc[line] and offsets[line + 1] (a piece with more
 *                   lines than its staging holds is scanned again, serially).
 *
 * Verdicts only (automata of at most 64 states; about twice as fast, see the
 * comment at k_text_verdicts below):
 *
 *   k_text_verdicts newlines numbered by the piece they lie in, counted with one
 *                   POPC per 16 bytes; the entry state of a piece that begins
 *                   inside a line is GUESSED and checked by the piece in front,
 *                   which leaves a correction when the guess was wrong.
 *   k_text_finish   chained scan of the counts, rc <- DECLINED / OK, corrections.
 *
 * Line i = buf[offsets[i], offsets[i+1]) includes its terminator, as for
 * sre_cuda_index_lines; a last line without '\n' ends at len.
 */
#include <cstdlib>

#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

constexpr uint32_t PIECE = 4096;
constexpr uint32_t MIN_PIECE = 512;     /* smallest piece the verdict path cuts (text_piece_bytes) */
constexpr uint32_t CAP = 64;            /* staged lines per piece */

/* bit 7 of every byte of x that equals '\n' (exact: no carry between bytes) */
__device__ __forceinline__ uint32_t nl_mask(uint32_t x)
{
    const uint32_t y = x ^ 0x0a0a0a0au;
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}

struct text_out_t {
    uint32_t       *stage;      /* [npieces][CAP] (end offset within the span << 1) | matched */
    uint32_t       *count;      /* [npieces] lines that start in the piece                  */
};

/* does a line start at the first byte of `piece`? */
__device__ __forceinline__ bool starts_line(const uint8_t *buf, size_t piece)
{
    return piece == 0 || __ldg(buf + piece * PIECE - 1) == '\n';
}

/*
 * Serial form over the table in global memory (the tail piece, and pieces with
 * more than CAP lines): the lines that START in [begin, end), each followed to
 * its end (at most `len`).  emit(end offset relative to begin, matched).
 */
template <class Emit>
__device__ __forceinline__ uint32_t serial_piece(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
                                                 size_t begin, size_t end, bool first_is_ours, Emit emit)
{
    const uint8_t *tx = dfa.x256;
    uint32_t s = dfa.start, n = 0;
    bool ours = first_is_ours;          /* the line that is open now started in [begin, end) */
    size_t p = begin;
    for (; p < len && (p < end || ours); p++) {
        const uint32_t b = __ldg(buf + p);
        s = __ldg(tx + ((s << 8) | b));
        if (b == '\n') {
            if (ours) {
                emit((uint32_t) (p + 1 - begin), s >> 7, n);
                n++;
            }
            ours = p + 1 < end;         /* where the next line starts */
        }
    }
    if (p == len && ours && len > begin && __ldg(buf + len - 1) != '\n') {
        /* a last line without terminator: the EOF step of the reference decides */
        const uint32_t st = s & 0x7f;
        emit((uint32_t) (len - begin), (st == dfa.acc || __ldg(dfa.fin + st)) ? 1u : 0u, n);
        n++;
    }
    return n;
}

/* MARK: the automaton has at most 64 states, and the table marks "this byte was a '\n'" in bit 6
 * of the state (rows r, r+64, r+128, r+192 are the same row): one OR + one test per word
 * replaces the newline search */
template <bool MARK>
struct text_consumer_t {
    const uint8_t      *tab;        /* x256 in shared memory */
    sre_dev_dfa_t       dfa;
    const uint8_t      *buf;
    size_t              len, npieces;
    text_out_t          out;
    uint32_t            s, pos;
    uint32_t            cnt;        /* lines recorded; 0xffffffff: the next line end is not ours */
    uint32_t           *stage;

    __device__ __forceinline__ void begin(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        s = dfa.start;
        pos = 0;
        cnt = 0xffffffffu;
        stage = out.stage;
        if (piece < npieces) {
            /* the first line end seen belongs to an earlier piece unless a line starts here */
            cnt = starts_line(buf, piece) ? 0u : 0xffffffffu;
            stage = out.stage + piece * CAP;
        }
    }
    __device__ __forceinline__ void record(uint32_t end_off, uint32_t matched)
    {
        if (cnt < CAP) {            /* (not for 0xffffffff: that line began in an earlier piece) */
            stage[cnt] = (end_off << 1) | matched;
        }
        cnt++;
    }
    __device__ __forceinline__ void word(uint32_t w, uint32_t at)
    {
        const uint32_t a0 = tab[__byte_perm(w, s, 0x5540)];
        const uint32_t a1 = tab[__byte_perm(w, a0, 0x5541)];
        const uint32_t a2 = tab[__byte_perm(w, a1, 0x5542)];
        const uint32_t a3 = tab[__byte_perm(w, a2, 0x5543)];
        s = a3;
        /* the states after each byte are still in registers: a word that holds a '\n' only adds
         * the bookkeeping */
        if (MARK) {
            if ((a0 | a1 | a2 | a3) & 0x40u) {
                if (a0 & 0x40u) record(at + 1, a0 >> 7);
                if (a1 & 0x40u) record(at + 2, a1 >> 7);
                if (a2 & 0x40u) record(at + 3, a2 >> 7);
                if (a3 & 0x40u) record(at + 4, a3 >> 7);
            }
        } else {
            const uint32_t m = nl_mask(w);
            if (m) {
                if (m & 0x00000080u) record(at + 1, a0 >> 7);
                if (m & 0x00008000u) record(at + 2, a1 >> 7);
                if (m & 0x00800000u) record(at + 3, a2 >> 7);
                if (m & 0x80000000u) record(at + 4, a3 >> 7);
            }
        }
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        word(v.x, pos);
        word(v.y, pos + 4);
        word(v.z, pos + 8);
        word(v.w, pos + 12);
        pos += 16;
    }
    __device__ __forceinline__ void byte(uint32_t b)
    {
        s = tab[(s << 8) | b];
        pos++;
        if (b == '\n') {
            record(pos, s >> 7);
        }
    }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        if (piece >= npieces) {
            return;
        }
        /* the line still open at the end of the piece (if it began here): read on, 16 aligned
         * bytes at a time (the piece ends on a 16-byte boundary), the next block requested
         * before this one is walked */
        const size_t base = piece * PIECE;
        size_t p = base + PIECE;
        if (cnt != 0xffffffffu && __ldg(buf + p - 1) != '\n') {
            bool done = false;
            uint4 vnext = make_uint4(0, 0, 0, 0);
            if (p < len) {
                vnext = __ldg(reinterpret_cast<const uint4 *>(buf + p));
            }
            while (p < len && !done) {
                const uint4 v = vnext;
                if (p + 16 < len) {
                    vnext = __ldg(reinterpret_cast<const uint4 *>(buf + p + 16));
                }
                const uint32_t w[4] = { v.x, v.y, v.z, v.w };
                const uint32_t n = len - p < 16 ? (uint32_t) (len - p) : 16u;
#pragma unroll
                for (uint32_t q = 0; q < 16; q++) {
                    if (q < n && !done) {
                        const uint32_t b = (w[q >> 2] >> ((q & 3) * 8)) & 0xff;
                        s = tab[(s << 8) | b];
                        if (b == '\n') {
                            record((uint32_t) (p + q + 1 - base), s >> 7);
                            done = true;
                        }
                    }
                }
                p += 16;
            }
            if (!done) {
                /* the buffer ended first: a last line without terminator (EOF step) */
                const uint32_t st = s & (MARK ? 0x3fu : 0x7fu);
                record((uint32_t) (len - base), (st == dfa.acc || __ldg(dfa.fin + st)) ? 1u : 0u);
            }
        }
        out.count[piece] = cnt == 0xffffffffu ? 0u : cnt;
    }
};

template <bool MARK>
__global__ void __launch_bounds__(1024, 1)
k_text_pieces(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ buf, size_t len,
              size_t npieces, text_out_t out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const dfa_smem_plan_t plan = dfa_smem_plan(256, 0, false);
    load_table(smem, MARK ? dfa.x256m : dfa.x256, 65536);
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    text_consumer_t<MARK> cons;
    cons.tab = smem;
    cons.dfa = dfa;
    cons.buf = buf;
    cons.len = len;
    cons.npieces = npieces;
    cons.out = out;
    tile_pipeline_tma_early<1>(cons, &tmap, npieces, PIECE,
                               smem + plan.stage_ofs + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + plan.bar_ofs) + warp * MAX_STAGES,
                               (size_t) warp * gridDim.x + blockIdx.x, (size_t) gridDim.x * warps_per_block);
}

/* the ragged tail piece [nfull * PIECE, len): one thread, table from global memory */
__global__ void k_text_tail(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len, size_t piece,
                            text_out_t out)
{
    uint32_t *stage = out.stage + piece * CAP;
    out.count[piece] = serial_piece(dfa, buf, len, piece * PIECE, len, starts_line(buf, piece),
                                    [&](uint32_t end_off, uint32_t matched, uint32_t k) {
                                        if (k < CAP) {
                                            stage[k] = (end_off << 1) | matched;
                                        }
                                    });
}

constexpr uint32_t WB = 1024;       /* pieces per block of the scan / write kernels */

/* block-wide exclusive prefix sum over WB threads; *total = the block's sum */
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[WB / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) {
            incl += up;
        }
    }
    if (lane == 31) {
        warp_sums[warp] = incl;
    }
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (uint32_t w = 0; w < WB / 32; w++) {
        const uint32_t x = warp_sums[w];
        before += w < warp ? x : 0;
        all += x;
    }
    __syncthreads();
    *total = all;
    return before + incl - v;
}

/* sums[b] <- lines that start in the WB pieces of block b */
__global__ void __launch_bounds__(WB)
k_text_sums(const uint32_t *__restrict__ count, size_t n, unsigned long long *__restrict__ sums, uint32_t mask)
{
    const size_t i = (size_t) blockIdx.x * WB + threadIdx.x;
    uint32_t total;
    block_exclusive(i < n ? (count[i] & mask) : 0u, &total);
    if (threadIdx.x == 0) {
        sums[blockIdx.x] = total;
    }
}

/* sums[b] <- lines before block b; sums[nb] <- all lines.  One block (nb = pieces / 1024). */
__global__ void __launch_bounds__(1024)
k_text_scan(unsigned long long *__restrict__ sums, size_t nb, unsigned long long *__restrict__ total_out)
{
    __shared__ unsigned long long carry, part[32];
    if (threadIdx.x == 0) {
        carry = 0;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t first = 0; first < nb; first += 1024) {
        const size_t i = first + threadIdx.x;
        const unsigned long long v = i < nb ? sums[i] : 0;
        unsigned long long incl = v;
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) {
                incl += up;
            }
        }
        if (lane == 31) {
            part[warp] = incl;
        }
        __syncthreads();
        unsigned long long before = 0, all = 0;
        for (uint32_t w = 0; w < 32; w++) {
            before += w < warp ? part[w] : 0;
            all += part[w];
        }
        const unsigned long long c = carry;
        if (i < nb) {
            sums[i] = c + before + incl - v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            carry = c + all;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sums[nb] = carry;
        *total_out = carry;
    }
}

/*
 * staging -> rc[] / offsets[].  A block takes WB pieces (one per thread for the
 * prefix sums); then every warp writes the lines of its 32 pieces together: lane
 * after lane takes the next line of the warp's run (the piece it belongs to found
 * by a search over the 32 prefix sums), so that consecutive lanes write
 * consecutive rows.
 */
__global__ void __launch_bounds__(WB)
k_text_write(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len, size_t npieces, text_out_t out,
             const unsigned long long *__restrict__ sums, int32_t *__restrict__ rc, int64_t *__restrict__ offsets,
             size_t max_lines)
{
    __shared__ uint32_t pre[WB];            /* lines of the block before each piece */
    const size_t piece = (size_t) blockIdx.x * WB + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, wfirst = threadIdx.x & ~31u;
    if (piece == 0 && offsets != nullptr) {
        offsets[0] = 0;
    }
    const uint32_t n = piece < npieces ? out.count[piece] : 0u;
    uint32_t total;
    const uint32_t before = block_exclusive(n, &total);
    pre[threadIdx.x] = before;
    __syncwarp();
    const size_t block_first = (size_t) sums[blockIdx.x];
    /* the warp's run of lines: [run0, run0 + run) of the block */
    const uint32_t run0 = __shfl_sync(FULL, before, 0);
    const uint32_t run = __shfl_sync(FULL, before + n, 31) - run0;
    const bool over = n > CAP;
    const bool any_over = __any_sync(FULL, over);
    for (uint32_t f = lane; f < run; f += 32) {
        /* the piece of the warp that line run0 + f belongs to: last one with pre <= run0 + f */
        uint32_t lo = 0, hi = 31;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (pre[wfirst + mid] <= run0 + f) {
                lo = mid;
            } else {
                hi = mid - 1;
            }
        }
        const uint32_t k = run0 + f - pre[wfirst + lo];
        const size_t pc = (size_t) blockIdx.x * WB + wfirst + lo, line = block_first + run0 + f;
        if (k < CAP && line < max_lines) {
            /* (rows of a piece with more than CAP lines are written again below) */
            const uint32_t r = out.stage[pc * CAP + k];
            rc[line] = (r & 1) ? SRE_K_OK : SRE_K_DECLINED;
            if (offsets != nullptr) {
                offsets[line + 1] = (int64_t) (pc * PIECE + (r >> 1));
            }
        }
    }
    if (any_over && over) {
        /* more lines than the staging holds: once more, serially, straight to the output */
        const size_t begin = piece * PIECE, end = begin + PIECE < len ? begin + PIECE : len;
        const size_t first = block_first + before;
        serial_piece(dfa, buf, len, begin, end, starts_line(buf, piece),
                     [&](uint32_t end_off, uint32_t matched, uint32_t k) {
                         const size_t line = first + k;
                         if (line < max_lines) {
                             rc[line] = matched ? SRE_K_OK : SRE_K_DECLINED;
                             if (offsets != nullptr) {
                                 offsets[line + 1] = (int64_t) (begin + end_off);
                             }
                         }
                     });
    }
}

/* ==== verdicts only, automata of at most 64 states: the count-only hot loop ==== */

/*
 * Every newline is numbered by the piece it lies in (line = newlines in the
 * pieces before + its index in the piece), so a thread never reads past its piece
 * and never tells "my" line ends from foreign ones.  The hot loop has no branch per
 * newline: the states after the four bytes of a word are packed into one register
 * (bit 6 = "this byte was a '\n'", bit 7 = "... and the line matched"), newlines
 * are counted with one POPC per 16 bytes and only a match bit (rare) leaves the
 * straight line.
 *
 * The line that is open where a piece begins was started by an earlier piece, so
 * the thread does not know the state it is entered with.  It GUESSES dfa.xguess
 * (the state the automaton idles in on ordinary text), and the thread of the piece
 * in front checks the guess when it gets there: exit state == guess -> the
 * neighbour's run IS the continuation and nothing is left to do; sticky ACC -> the
 * line matched whatever follows; anything else (a partial match hanging over the
 * boundary) -> it finishes the line itself, byte by byte, and leaves the verdict as
 * a correction in its info word, which k_text_finish applies to the first line of
 * the next piece that holds a newline.  A correct guess is a matter of speed only.
 */
constexpr uint32_t INFO_CNT = 0x1fffu;          /* newlines in the piece (<= 4097)            */
constexpr uint32_t INFO_MSHIFT = 13;            /* matched ones among them                     */
constexpr uint32_t INFO_FSHIFT = 26;            /* verdict of the line open at the piece's end */
constexpr uint32_t INFO_BEGINS = 1u << 28;      /* a line begins at the first byte of the piece  */
constexpr uint32_t FIX_NONE = 0, FIX_UNMATCHED = 1, FIX_MATCHED = 2;
constexpr uint32_t NLBITS = 0x40404040u, MBITS = 0x80808080u;

struct verdict_out_t {
    uint32_t *info;         /* [nfull + 1] */
    uint32_t *stage;        /* [nfull + 1][CAP] piece-local numbers of the matched newlines */
};

struct verdict_consumer_t {
    uint32_t            tab_s;      /* x256m in shared memory, rows padded to ROW260 bytes (row v starts v
                                       banks further on: lanes in different states do not collide on the
                                       bank of the byte they read); shared-window address, low byte 0 */
    const uint8_t      *buf, *fin;
    size_t              len, nfull;
    verdict_out_t       out;
    uint32_t            piece_bytes, start, guess, acc;
    uint32_t            s, cnt, mcnt, begins;
    uint32_t           *stage;

    __device__ __forceinline__ void begin(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        cnt = 0;
        mcnt = 0;
        begins = 0;
        s = guess;
        stage = out.stage;
        if (piece < nfull) {
            begins = (piece == 0 || __ldg(buf + piece * piece_bytes - 1) == '\n') ? INFO_BEGINS : 0u;
            s = begins ? start : guess;
            stage = out.stage + piece * CAP;
        }
    }
    __device__ __forceinline__ uint32_t look(uint32_t state, uint32_t b) const
    {
        return step260_t::lds_u8(state * ROW260 + tab_s + b);
    }
    /* the four bytes of w; -> the states after each of them, one per byte.  The byte's address in
     * row 0 (PRMT, off the chain) leaves one IMAD + one LDS per byte on the state -> state chain. */
    __device__ __forceinline__ uint32_t word(uint32_t w)
    {
        const uint32_t b0 = __byte_perm(w, tab_s, 0x7650), b1 = __byte_perm(w, tab_s, 0x7651);
        const uint32_t b2 = __byte_perm(w, tab_s, 0x7652), b3 = __byte_perm(w, tab_s, 0x7653);
        const uint32_t a0 = step260_t::lds_u8(s * ROW260 + b0);
        const uint32_t a1 = step260_t::lds_u8(a0 * ROW260 + b1);
        const uint32_t a2 = step260_t::lds_u8(a1 * ROW260 + b2);
        const uint32_t a3 = step260_t::lds_u8(a2 * ROW260 + b3);
        s = a3;
        return (a3 * 256u + a2) * 65536u + (a1 * 256u + a0);
    }
    /* some line that ends in these 16 bytes matched: stage the numbers of those newlines */
    __device__ __forceinline__ void matches(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3)
    {
        const uint32_t p[4] = { p0, p1, p2, p3 };
        uint32_t k = cnt;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t e = p[i] >> (8 * q);
                if (e & 0x40u) {
                    if (e & 0x80u) {
                        if (mcnt < CAP) {
                            stage[mcnt] = k;
                        }
                        mcnt++;
                    }
                    k++;
                }
            }
        }
    }
    __device__ __forceinline__ void chunk(const uint4 &v)
    {
        const uint32_t p0 = word(v.x), p1 = word(v.y), p2 = word(v.z), p3 = word(v.w);
        if ((p0 | p1 | p2 | p3) & MBITS) {
            matches(p0, p1, p2, p3);
        }
        /* the sixteen newline bits side by side, one POPC */
        const uint32_t nl = (p0 & NLBITS) | ((p1 >> 1) & (NLBITS >> 1)) | ((p2 >> 2) & (NLBITS >> 2))
                            | ((p3 >> 3) & (NLBITS >> 3));
        cnt += __popc(nl);
    }
    __device__ __forceinline__ void byte(uint32_t b)
    {
        s = look(s, b);
        if (s & 0x40u) {
            if (s & 0x80u) {
                if (mcnt < CAP) {
                    stage[mcnt] = cnt;
                }
                mcnt++;
            }
            cnt++;
        }
    }
    __device__ __forceinline__ void end(size_t group)
    {
        const size_t piece = group * 32 + (threadIdx.x & 31);
        if (piece >= nfull) {
            return;
        }
        uint32_t fix = FIX_NONE;
        const uint32_t st = s & 0x3fu;
        if (!(s & 0x40u) && st != guess) {
            /* a line is open here and the next piece's thread did not start from this state */
            if (st == acc) {
                fix = FIX_MATCHED;
            } else {
                /* a partial match hangs over the boundary: finish the line -- 16 aligned bytes at
                 * a time (pieces end on 128-byte boundaries), the next block requested before this
                 * one is walked */
                size_t p = (piece + 1) * piece_bytes;
                uint32_t t = s;
                bool done = false;
                uint4 vnext = make_uint4(0, 0, 0, 0);
                if (p + 16 <= len) {
                    vnext = __ldg(reinterpret_cast<const uint4 *>(buf + p));
                }
                while (p + 16 <= len && !done) {
                    const uint4 v = vnext;
                    if (p + 32 <= len) {
                        vnext = __ldg(reinterpret_cast<const uint4 *>(buf + p + 16));
                    }
                    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                    for (uint32_t q = 0; q < 16; q++) {
                        if (!done) {
                            t = look(t, (w[q >> 2] >> ((q & 3) * 8)) & 0xff);
                            done = (t & 0x40u) != 0;
                        }
                    }
                    p += 16;
                }
                for (; p < len && !done; p++) {
                    t = look(t, __ldg(buf + p));
                    done = (t & 0x40u) != 0;
                }
                if (done) {
                    fix = (t & 0x80u) ? FIX_MATCHED : FIX_UNMATCHED;
                } else {
                    /* the buffer ended first: the EOF step of the reference decides */
                    const uint32_t e = t & 0x3fu;
                    fix = (e == acc || __ldg(fin + e)) ? FIX_MATCHED : FIX_UNMATCHED;
                }
            }
        }
        out.info[piece] = cnt | (mcnt << INFO_MSHIFT) | (fix << INFO_FSHIFT) | begins;
    }
};

/*
 * The same walk, serially, over the table in global memory: the newlines of
 * [begin, end), emit(piece-local number) for those that end a matched line.  With
 * `eof` (the tail piece) a last line without terminator counts as one more newline,
 * its verdict the reference's EOF step.  -> newlines counted
 */
template <class Emit>
__device__ __forceinline__ uint32_t serial_marks(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len, size_t begin,
                                                 size_t end, bool eof, Emit emit)
{
    const uint8_t *tx = dfa.x256m;
    uint32_t s = (begin == 0 || __ldg(buf + begin - 1) == '\n') ? dfa.start : dfa.xguess, n = 0;
    auto step = [&](uint32_t b) {
        s = __ldg(tx + ((s << 8) | b));
        if (s & 0x40u) {
            if (s & 0x80u) {
                emit(n);
            }
            n++;
        }
    };
    /* 16 aligned bytes at a time (pieces begin on 128-byte boundaries), the next block requested
     * before this one is walked: the table look-ups (L1) are the chain, not the input */
    size_t p = begin;
    uint4 vnext = make_uint4(0, 0, 0, 0);
    if (p + 16 <= end) {
        vnext = __ldg(reinterpret_cast<const uint4 *>(buf + p));
    }
    while (p + 16 <= end) {
        const uint4 v = vnext;
        if (p + 32 <= end) {
            vnext = __ldg(reinterpret_cast<const uint4 *>(buf + p + 16));
        }
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (uint32_t q = 0; q < 16; q++) {
            step((w[q >> 2] >> ((q & 3) * 8)) & 0xff);
        }
        p += 16;
    }
    for (; p < end; p++) {
        step(__ldg(buf + p));
    }
    if (eof && len > 0 && __ldg(buf + len - 1) != '\n') {
        const uint32_t st = s & 0x3fu;
        if (st == dfa.acc || __ldg(dfa.fin + st)) {
            emit(n);
        }
        n++;
    }
    return n;
}

/* the tail piece [nfull * piece_bytes, len) (possibly empty) and the end of the buffer: one thread */
__device__ __forceinline__ void text_tail_piece(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len, size_t nfull,
                                                uint32_t piece_bytes, const verdict_out_t &out)
{
    uint32_t *stage = out.stage + nfull * CAP;
    uint32_t m = 0;
    const size_t begin = nfull * piece_bytes;
    const uint32_t n = serial_marks(dfa, buf, len, begin, len, true, [&](uint32_t k) {
        if (m < CAP) {
            stage[m] = k;
        }
        m++;
    });
    const uint32_t begins = (begin == 0 || __ldg(buf + begin - 1) == '\n') ? INFO_BEGINS : 0u;
    out.info[nfull] = n | (m << INFO_MSHIFT) | begins;
}

/* shared memory: [table 256 x ROW260][barriers][stages] */
constexpr size_t VTAB_BYTES = align_up((size_t) 256 * ROW260, 1024);
constexpr size_t VBAR_OFS = VTAB_BYTES, VSTAGE_OFS = VTAB_BYTES + 2048;

__global__ void __launch_bounds__(1024, 1)
k_text_verdicts(sre_dev_dfa_t dfa, const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ buf, size_t len,
                size_t nfull, uint32_t piece_bytes, verdict_out_t out, unsigned long long *__restrict__ status,
                size_t nstatus)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    load_table260(smem, dfa.x256m, 256);
    if (blockIdx.x == 0) {
        /* the ticket and the look-back words of k_text_finish */
        for (size_t i = threadIdx.x; i < nstatus; i += blockDim.x) {
            status[i] = 0;
        }
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    verdict_consumer_t cons;
    cons.tab_s = smem_u32(smem);
    cons.buf = buf;
    cons.fin = dfa.fin;
    cons.len = len;
    cons.nfull = nfull;
    cons.out = out;
    cons.piece_bytes = piece_bytes;
    cons.start = dfa.start;
    cons.guess = dfa.xguess;
    cons.acc = dfa.acc;
    tile_pipeline_tma_early<1>(cons, &tmap, nfull, piece_bytes, smem + VSTAGE_OFS + (size_t) warp * 32 * 128,
                               reinterpret_cast<uint64_t *>(smem + VBAR_OFS) + warp * MAX_STAGES,
                               (size_t) warp * gridDim.x + blockIdx.x, (size_t) gridDim.x * warps_per_block);
    /* the tail piece (shorter than a piece, not a row of the tensor): one lane of the warp that
     * was handed the fewest pieces walks it when its own are done */
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 32) {
        text_tail_piece(dfa, buf, len, nfull, piece_bytes, out);
    }
}

/* inputs shorter than a piece: the tail piece and the ticket + look-back words of k_text_finish
 * (otherwise k_text_verdicts does both) */
__global__ void __launch_bounds__(256)
k_text_verdicts_tail(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len, size_t nfull,
                     uint32_t piece_bytes, verdict_out_t out, unsigned long long *__restrict__ status, size_t nstatus)
{
    for (size_t i = threadIdx.x; i < nstatus; i += blockDim.x) {
        status[i] = 0;
    }
    if (threadIdx.x == 0) {
        text_tail_piece(dfa, buf, len, nfull, piece_bytes, out);
    }
}

/*
 * Everything after the pieces in one launch: line numbers (a chained scan over the blocks: each
 * block publishes its count, then the count of all blocks up to it; a block looks back until it
 * meets such a total -- tiles are handed out by ticket, so the blocks it waits for are running),
 * rc <- SRE_DECLINED for the block's lines, SRE_OK for the staged matched ones, then the
 * corrections.  One thread per piece.  status[0] = ticket, status[1 + b] = (count << 2) | 1 (the
 * block's own lines) or | 2 (all lines up to and including block b).
 */
__global__ void __launch_bounds__(WB)
k_text_finish(sre_dev_dfa_t dfa, const uint8_t *__restrict__ buf, size_t len, size_t nfull, uint32_t piece_bytes,
              verdict_out_t out, unsigned long long *status, unsigned long long *__restrict__ total_out,
              int32_t *__restrict__ rc, size_t max_lines)
{
    __shared__ unsigned long long s_first;
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) {
        s_tile = (uint32_t) atomicAdd(&status[0], 1ull);
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t piece = (size_t) tile * WB + threadIdx.x;
    const uint32_t info = piece <= nfull ? out.info[piece] : 0u;
    const uint32_t n = info & INFO_CNT;
    uint32_t total;
    const uint32_t before = block_exclusive(n, &total);
    if (threadIdx.x < 32) {
        /* warp 0 looks back 32 blocks at a time */
        volatile unsigned long long *st = status + 1;
        const uint32_t lane = threadIdx.x;
        unsigned long long excl = 0;
        if (tile != 0) {
            if (lane == 0) {
                st[tile] = ((unsigned long long) total << 2) | 1ull;
                __threadfence();
            }
            for (int64_t hi = (int64_t) tile - 1; hi >= 0; hi -= 32) {
                const int64_t j = hi - lane;            /* lane 0 = the nearest block */
                unsigned long long v = 2ull;            /* (before block 0: a total of 0) */
                if (j >= 0) {
                    while (((v = st[j]) & 3ull) == 0) {
                    }
                }
                /* the nearest block that knows the total up to itself ends the walk */
                const uint32_t done = __ballot_sync(FULL, (v & 3ull) == 2ull);
                const uint32_t upto = done ? (uint32_t) __ffs((int) done) - 1 : 31u;
                unsigned long long part = lane <= upto ? v >> 2 : 0ull;
#pragma unroll
                for (uint32_t d = 16; d > 0; d >>= 1) {
                    part += __shfl_xor_sync(FULL, part, d);
                }
                excl += part;
                if (done) {
                    break;
                }
            }
        }
        if (lane == 0) {
            st[tile] = ((excl + total) << 2) | 2ull;
            __threadfence();
            s_first = excl;
            if ((size_t) tile == nfull / WB) {      /* the block that holds the tail piece */
                *total_out = excl + total;
            }
        }
    }
    __syncthreads();
    const size_t block_first = (size_t) s_first;
    for (size_t i = threadIdx.x; i < total; i += WB) {
        if (block_first + i < max_lines) {
            rc[block_first + i] = SRE_K_DECLINED;
        }
    }
    __syncthreads();
    if (piece > nfull || n == 0) {
        return;
    }
    const size_t first = block_first + before;
    const uint32_t m = (info >> INFO_MSHIFT) & INFO_CNT;
    if (m <= CAP) {
        const uint32_t *stage = out.stage + piece * CAP;
        for (uint32_t k = 0; k < m; k++) {
            const size_t line = first + stage[k];
            if (line < max_lines) {
                rc[line] = SRE_K_OK;
            }
        }
    } else {
        const size_t begin = piece * piece_bytes, end = piece < nfull ? begin + piece_bytes : len;
        serial_marks(dfa, buf, len, begin, end, piece == nfull, [&](uint32_t k) {
            if (first + k < max_lines) {
                rc[first + k] = SRE_K_OK;
            }
        });
    }
    /* my first newline ends a line that an earlier piece began: the first correction left by the
     * pieces from its owner on (the owner = the nearest piece before me that holds a newline or
     * begins a line; the pieces between are all inside the line) decides; none: the guess held */
    if (piece > 0 && first < max_lines && !(info & INFO_BEGINS)) {
        size_t k = piece - 1;
        while (k > 0 && (out.info[k] & (INFO_CNT | INFO_BEGINS)) == 0) {
            k--;
        }
        for (; k < piece; k++) {
            const uint32_t fix = (out.info[k] >> INFO_FSHIFT) & 3u;
            if (fix != FIX_NONE) {
                rc[first] = fix == FIX_MATCHED ? SRE_K_OK : SRE_K_DECLINED;
                break;
            }
        }
    }
}

}  // namespace

/*
 * The piece size of the verdict path.  A thread walks its piece serially (~40 cycles per byte),
 * so below ~600 MB -- one 4 KB piece per resident thread -- the call takes the time of ONE piece
 * however small the input: pieces are cut so that every resident thread gets one (measured:
 * 64 MiB in 4 KB pieces 0.18 ms, in 1 KB pieces 0.08 ms), down to MIN_PIECE.  Above that the
 * kernel is bound by the shared-memory pipe, not by latency, and a thinner last round simply runs
 * faster (1 GiB = 1.73 rounds of 4 KB pieces: 3.5 TB/s).  SRE_CUDA_TEXT_PIECE overrides (tuning /
 * tests): a multiple of 128.
 */
static uint32_t text_piece_bytes(size_t len, size_t threads_total)
{
    const char *e = getenv("SRE_CUDA_TEXT_PIECE");
    if (e != nullptr) {
        const long v = atol(e);
        if (v >= (long) MIN_PIECE && v <= (long) PIECE && v % 128 == 0) {
            return (uint32_t) v;
        }
    }
    /* the smallest power of two that gives every resident thread at most one piece.  (Powers of
     * two only: pieces of 7, 9, 14, 28 or 31 tiles were measured up to 1.8x slower than the next
     * power of two at the same input size -- 512 MiB: 3584 bytes 0.31 ms, 4096 bytes 0.20 ms.) */
    uint32_t per = MIN_PIECE;
    while (per < PIECE && (size_t) per * threads_total < len) {
        per *= 2;
    }
    return per;
}

size_t sre_text_workspace_bytes(size_t len)
{
    /* whichever path runs: the pieces the verdict path would cut, or 4 KB pieces */
    const uint32_t piece = text_piece_bytes(len, (size_t) num_sms() * 1024);
    const size_t npieces = len / piece + 2, nb = (npieces + WB - 1) / WB;
    return 256 + (nb + 2) * 8 + npieces * (CAP * 4 + 8) + 2048;
}

/* where in the workspace the number of lines is left (8 bytes) */
size_t sre_text_count_offset(size_t)
{
    return 0;
}

/* verdicts only, <= 64 states: count-only hot loop (k_text_verdicts) */
static cudaError_t launch_text_verdicts(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len, int32_t *rc,
    size_t max_lines, uint8_t *workspace, cudaStream_t stream, int *launches)
{
    const int warps = 32;
    const uint32_t piece = text_piece_bytes(len, (size_t) num_sms() * warps * 32);
    const size_t nfull = len / piece, npieces = nfull + 1, nb = (npieces + WB - 1) / WB;
    unsigned long long *total = reinterpret_cast<unsigned long long *>(workspace);
    unsigned long long *sums = reinterpret_cast<unsigned long long *>(workspace + 256);     /* [nb + 1] */
    uint8_t *p = workspace + 256 + ((nb + 2) * 8 + 255) / 256 * 256;
    verdict_out_t out;
    out.info = reinterpret_cast<uint32_t *>(p);
    p += (npieces * 4 + 255) / 256 * 256;
    out.stage = reinterpret_cast<uint32_t *>(p);
    cudaError_t err;
    if (nfull) {
        const size_t smem = VSTAGE_OFS + (size_t) warps * 32 * 128;
        CUtensorMap tmap;
        if ((err = make_row_tensor_map(&tmap, buf, nfull, piece, 128)) != cudaSuccess) return err;
        static bool attr_set = false;
        if (!attr_set) {
            err = cudaFuncSetAttribute(k_text_verdicts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (err != cudaSuccess) return err;
            attr_set = true;
        }
        const size_t ngroups = (nfull + 31) / 32;
        size_t grid = (size_t) num_sms();
        const size_t need = (ngroups + warps - 1) / warps;
        if (grid > need) {
            grid = need;
        }
        if (launches) ++*launches;
        /* (sums[] serves as the ticket + look-back words of k_text_finish) */
        k_text_verdicts<<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, tmap, buf, len, nfull, piece, out, sums,
                                                                       nb + 1);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    } else {
        if (launches) ++*launches;
        k_text_verdicts_tail<<<1, 256, 0, stream>>>(dfa, buf, len, nfull, piece, out, sums, nb + 1);
    }
    if (launches) ++*launches;
    k_text_finish<<<(unsigned) nb, WB, 0, stream>>>(dfa, buf, len, nfull, piece, out, sums, total, rc, max_lines);
    return cudaGetLastError();
}

/* workspace: sre_text_workspace_bytes(len) bytes, 256-byte aligned */
cudaError_t sre_launch_text(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len, int32_t *rc,
    int64_t *offsets, size_t max_lines, uint8_t *workspace, cudaStream_t stream, int *launches)
{
    if (dfa.x256 == nullptr || (reinterpret_cast<uintptr_t>(buf) & 15)) {
        return cudaErrorInvalidValue;
    }
    if (offsets == nullptr && dfa.x256m != nullptr) {
        return launch_text_verdicts(dfa, buf, len, rc, max_lines, workspace, stream, launches);
    }
    const size_t nfull = len / PIECE, npieces = nfull + (len % PIECE ? 1 : 0), nb = (npieces + WB - 1) / WB;
    unsigned long long *total = reinterpret_cast<unsigned long long *>(workspace);
    unsigned long long *sums = reinterpret_cast<unsigned long long *>(workspace + 256);     /* [nb + 1] */
    uint8_t *p = workspace + 256 + ((nb + 2) * 8 + 255) / 256 * 256;
    text_out_t out;
    out.count = reinterpret_cast<uint32_t *>(p);
    p += (npieces * 4 + 255) / 256 * 256 + 256;
    out.stage = reinterpret_cast<uint32_t *>(p);
    cudaError_t err;
    if (npieces == 0) {
        if ((err = cudaMemsetAsync(workspace, 0, 8, stream)) != cudaSuccess) return err;
        return offsets ? cudaMemsetAsync(offsets, 0, 8, stream) : cudaSuccess;
    }
    if (nfull) {
        const dfa_smem_plan_t plan = dfa_smem_plan(256, 0, false);
        const int warps = 32;
        const size_t smem = plan.stage_ofs + (size_t) warps * 32 * 128;
        CUtensorMap tmap;
        if ((err = make_row_tensor_map(&tmap, buf, nfull, PIECE, 128)) != cudaSuccess) return err;
        static bool attr_set = false;
        if (!attr_set) {
            err = cudaFuncSetAttribute(k_text_pieces<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (err == cudaSuccess) {
                err = cudaFuncSetAttribute(k_text_pieces<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int) smem);
            }
            if (err != cudaSuccess) return err;
            attr_set = true;
        }
        const size_t ngroups = (nfull + 31) / 32;
        size_t grid = (size_t) num_sms();
        const size_t need = (ngroups + warps - 1) / warps;
        if (grid > need) {
            grid = need;
        }
        if (launches) ++*launches;
        if (dfa.x256m != nullptr) {
            k_text_pieces<true><<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, tmap, buf, len, nfull, out);
        } else {
            k_text_pieces<false><<<(unsigned) grid, warps * 32, smem, stream>>>(dfa, tmap, buf, len, nfull, out);
        }
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    if (npieces > nfull) {
        if (launches) ++*launches;
        k_text_tail<<<1, 1, 0, stream>>>(dfa, buf, len, nfull, out);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    if (launches) *launches += 3;
    k_text_sums<<<(unsigned) nb, WB, 0, stream>>>(out.count, npieces, sums, 0xffffffffu);
    k_text_scan<<<1, 1024, 0, stream>>>(sums, nb, total);
    k_text_write<<<(unsigned) nb, WB, 0, stream>>>(dfa, buf, len, npieces, out, sums, rc, offsets, max_lines);
    return cudaGetLastError();
}
