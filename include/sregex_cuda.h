/*
 * sregex_cuda.h -- batch / device-pointer extension of the sregex C API.
 *
 * New in this build.  The reference's executors take one buffer per call
 * (sre_vm_thompson_exec, sregex.h:147-148; sre_vm_pike_exec, sregex.h:133-134;
 * the JIT handler type, sregex.h:158-159); a kernel launch per 1 KB line would
 * cap throughput at launch rate, so the GPU library adds entry points that take
 * a whole corpus.  Each function below names the reference call it is the batch
 * form of.  Plain C ABI: pointers and sizes only.  Pointers named dev_* are CUDA
 * device pointers; `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Functions returning int give SRE_OK or SRE_ERROR unless noted; there
 * is no CPU fallback -- without a usable CUDA device they return SRE_ERROR.
 *
 * Threading: a lowered program is immutable.  Every call takes its device
 * scratch from the device's stream-ordered memory pool on the caller's stream
 * and returns it there, so one program may be used from any number of host
 * threads and CUDA streams at the same time (the reference's JIT run time is
 * re-entrant in the same way; its interpreter and Pike VM are not, since they
 * write dedup tags into the program: sre_vm_thompson.c:284, sre_vm_pike.c:792).
 * The host-buffer (*_host) and classic (sregex.h) entry points work on the
 * calling thread's per-thread default stream.
 */
#ifndef SREGEX_B200_SREGEX_CUDA_H
#define SREGEX_B200_SREGEX_CUDA_H

#include <sregex/sregex.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GPU tables lowered from a compiled program; cached inside the program and
 * released with the pool that owns the program. */
typedef struct sre_cuda_program_s  sre_cuda_program_t;

enum {
    SRE_CUDA_ENGINE_AUTO        = 0,    /* best tier available                       */
    SRE_CUDA_ENGINE_DFA_TILED   = 1,    /* smem-staged thread-per-line DFA           */
    SRE_CUDA_ENGINE_DFA_GENERIC = 2,    /* thread-per-line DFA, any alignment        */
    SRE_CUDA_ENGINE_NFA         = 3,    /* bit-parallel NFA: thread per line for <=
                                           64 lowered states (aligned lines), else
                                           warp per line                             */
    SRE_CUDA_ENGINE_DFA_SKIP    = 4,    /* DFA_TILED + word skip; needs a start state
                                           left by <= 4 byte values (AUTO picks it
                                           when 1 or 2 byte values leave it)          */
    SRE_CUDA_ENGINE_NFA_WARP    = 5     /* warp-per-line bit-parallel NFA (any size
                                           up to 4096 lowered states)                */
};
/* tuning: launch shape of DFA_TILED / DFA_SKIP (0 = default); pass
 * engine | SRE_CUDA_ENGINE_VARIANT(v) */
#define SRE_CUDA_ENGINE_VARIANT(v)  (((v) & 0xff) << 8)

typedef struct {
    uint32_t  prog_len;         /* bytecode instructions                            */
    uint32_t  nfa_states;       /* lowered (pc, look-ahead) states                  */
    uint32_t  nfa_classes;      /* byte classes                                     */
    uint32_t  nfa_kinds;        /* 3 if look-behind assertions are present          */
    uint32_t  nfa_shift_states; /* states handled by the shift path                 */
    uint32_t  dfa_states;       /* 0: subset construction gave up                   */
    uint32_t  dfa_classes;
    uint32_t  dfa_byte_table;   /* 1: [state][byte] u8 table (<= 256 states)        */
    uint32_t  dfa_leave_bytes;  /* byte values leaving the start state (1..4), else 0 */
    uint32_t  nregexes;
    uint32_t  pike_slots;       /* capture slots of the whole set                   */
    uint64_t  pike_ctx_bytes;   /* device scratch per concurrent Pike context       */
    uint32_t  dfa_start;        /* DFA state a stream begins in                     */
    uint32_t  dfa_acc;          /* absorbing "a step saw a live MATCH thread" state */
    uint32_t  image_states;     /* states of the stream scan's image automaton      */
    uint32_t  reserved;
} sre_cuda_info_t;

/* Lowering pass (host) + upload.  The batch analogue of
 * sre_vm_thompson_jit_compile (sregex.h:162-163): done once per program. */
SRE_API sre_cuda_program_t *sre_cuda_program_create(sre_program_t *prog);
SRE_API int sre_cuda_program_info(sre_cuda_program_t *cp, sre_cuda_info_t *info);

/*
 * Batch form of sre_vm_thompson_exec(ctx, line, len, eof=1) with a fresh ctx per
 * line (how bench/sregex.c:224-228 and the nginx module use it).  Line i is
 * dev_buf[i*pitch, i*pitch + linelen).  dev_rc[i] = SRE_OK or SRE_DECLINED.
 */
SRE_API int sre_cuda_thompson_exec_lines(sre_cuda_program_t *cp,
    const uint8_t *dev_buf, size_t nlines, size_t pitch, size_t linelen,
    int32_t *dev_rc, int engine, void *stream);

/* Same with ragged lines: line i is dev_buf[dev_offsets[i], dev_offsets[i+1]). */
SRE_API int sre_cuda_thompson_exec_ragged(sre_cuda_program_t *cp,
    const uint8_t *dev_buf, const int64_t *dev_offsets, size_t nlines,
    int32_t *dev_rc, int engine, void *stream);

/*
 * Batch form of sre_vm_pike_exec(ctx, line, len, eof=1, NULL) with a fresh ctx
 * per line.  dev_rc[i] = matched regex id (>= 0), SRE_DECLINED or SRE_ERROR;
 * dev_ovec[i*ovec_slots ..] = what the reference leaves in the caller's ovector
 * (matched regex's groups, -1 fill; all -1 when there is no match).
 * dev_offsets may be NULL (fixed pitch); with dev_offsets, pitch is ignored and
 * linelen, when non-zero, is an upper bound on the line lengths (below 32 KB it
 * lets the kernels keep capture offsets in 16 bits).  dev_select: when given, only lines with
 * dev_select[i] == SRE_OK are run (others: rc = dev_select[i]), which lets a
 * Thompson pass gate the capture pass.  When dev_select is NULL and the lines
 * are aligned fixed-pitch, the library runs that gate itself with the
 * determinised program, which also yields per line the offset after which no
 * earlier-started thread is alive; the Pike search of a line starts there (an
 * exact form of the reference's first-byte prefilter, sre_vm_pike.c:256-309).
 */
SRE_API int sre_cuda_pike_exec_lines(sre_cuda_program_t *cp,
    const uint8_t *dev_buf, const int64_t *dev_offsets, size_t nlines,
    size_t pitch, size_t linelen, const int32_t *dev_select, int32_t *dev_rc,
    int64_t *dev_ovec, size_t ovec_slots, void *stream);

/*
 * Global scan: all non-overlapping matches of every line, i.e. the batch form of
 * calling sre_vm_pike_exec again after each match on the rest of the data (the
 * post-match continuation of sre_vm_pike.c:624-635 incl. the one-byte skip after
 * an empty match, :179-193 -- what ngx_replace_filter does).  Per line at most
 * max_matches matches: dev_count[i], dev_spans[i][k] = ($0 start, $0 end),
 * dev_ids[i][k] = regex id.
 */
SRE_API int sre_cuda_pike_exec_lines_all(sre_cuda_program_t *cp,
    const uint8_t *dev_buf, const int64_t *dev_offsets, size_t nlines,
    size_t pitch, size_t linelen, size_t max_matches, int32_t *dev_count,
    int64_t *dev_spans, int32_t *dev_ids, void *stream);

/*
 * Many concurrent streams on the Pike VM -- the ngx_replace_filter shape: one
 * sre_vm_pike_ctx_t per connection, fed chunk by chunk with SRE_AGAIN, temp
 * captures and pending matches (sre_vm_pike.c:148-689, :624-658, :692-735).
 * _create makes nstreams persistent contexts on the device; every _exec call is
 * one sre_vm_pike_exec(ctx_i, chunk_i, len_i, eof_i, &pending) per stream,
 * chunk_i = dev_buf[dev_offsets[i], dev_offsets[i+1]) (may be empty), eof_i =
 * eof_all || dev_eof[i] (dev_eof may be NULL).  Row i of dev_out (4 + ovec_slots
 * values): rc; 1 if a pending match is reported; its span ($0 start, end); then
 * the ovector as the reference leaves it (whole on a match, slots 0-1 = the temp
 * capture on SRE_AGAIN).  A stream that has returned a match continues after it
 * on its next call, like the reference's ctx.
 */
typedef struct sre_cuda_pike_streams_s  sre_cuda_pike_streams_t;
SRE_API sre_cuda_pike_streams_t *sre_cuda_pike_streams_create(sre_cuda_program_t *cp,
    size_t nstreams, void *stream);
SRE_API int sre_cuda_pike_streams_exec(sre_cuda_pike_streams_t *streams,
    const uint8_t *dev_buf, const int64_t *dev_offsets, const uint8_t *dev_eof,
    unsigned eof_all, int64_t *dev_out, size_t ovec_slots, void *stream);
SRE_API void sre_cuda_pike_streams_free(sre_cuda_pike_streams_t *streams);

/*
 * Chunk-parallel form of a sequence of sre_vm_thompson_exec(ctx, chunk_k,
 * chunk_bytes, eof) calls over one long stream resident on the device
 * (sre_vm_thompson.c:63-270; the carried thread lists are one DFA state here).
 * Works for every program that has a DFA (any number of states).
 * *state_io carries the automaton state across calls (set it to
 * SRE_CUDA_STATE_INIT before the first call).  Returns SRE_OK / SRE_AGAIN /
 * SRE_DECLINED like the reference would after the last chunk; on SRE_OK
 * *match_chunk (if not NULL) is the index of the chunk_bytes-sized chunk in
 * which the reference's call sequence first returns SRE_OK.  dev_buf must be
 * 16-byte aligned.
 */
#define SRE_CUDA_STATE_INIT     0xffffffffu     /* the state a stream begins in       */
#define SRE_CUDA_STATE_UNKNOWN  0xfffffffeu     /* entry state not known (yet)        */
SRE_API int sre_cuda_thompson_exec_stream(sre_cuda_program_t *cp,
    const uint8_t *dev_buf, size_t len, size_t chunk_bytes, unsigned eof,
    uint32_t *state_io, int64_t *match_chunk, void *stream);

/*
 * Pieces of the stream scan, exposed for multi-GPU sharding: every
 * rank owns a contiguous part of the stream.
 *   1. the ranks exchange halos: the last SRE_CUDA_STREAM_HALO bytes of each part;
 *   2. _reduce turns a part into one record (SRE_CUDA_STREAM_FN_BYTES bytes,
 *      host): the states the part can be entered in -- whatever the stream
 *      held before the halo -- and the state each of them leads to.  dev_halo =
 *      the preceding part's halo (device), or NULL with entry_state given (the
 *      first part: SRE_CUDA_STATE_INIT) or SRE_CUDA_STATE_UNKNOWN;
 *   3. the records are all-gathered; sre_cuda_stream_fn_apply chains them in
 *      rank order from the stream's start state, which gives every rank its
 *      true entry state (SRE_CUDA_STATE_UNKNOWN from a record that is still
 *      unresolved: its rank must _resolve first and publish the new record);
 *   4. _resolve yields the part's exit state and the offset (within the part)
 *      of the step that first sees a match, -1 if none, and the part's record
 *      again, now {entry -> exit}.
 * The scan handle keeps the part's records on the device between 2 and 4.
 */
#define SRE_CUDA_STREAM_FN_BYTES  32
#define SRE_CUDA_STREAM_HALO      256
typedef struct sre_cuda_stream_scan_s  sre_cuda_stream_scan_t;
SRE_API sre_cuda_stream_scan_t *sre_cuda_thompson_stream_reduce(sre_cuda_program_t *cp,
    const uint8_t *dev_buf, size_t len, const uint8_t *dev_halo, uint32_t entry_state,
    uint8_t *host_fn /* [SRE_CUDA_STREAM_FN_BYTES], may be NULL */, void *stream);
SRE_API int sre_cuda_thompson_stream_resolve(sre_cuda_stream_scan_t *scan,
    uint32_t entry_state, uint32_t *exit_state, int64_t *first_match_offset,
    uint8_t *host_fn /* may be NULL */);
SRE_API void sre_cuda_thompson_stream_free(sre_cuda_stream_scan_t *scan);
SRE_API uint32_t sre_cuda_stream_fn_apply(const uint8_t *fn, uint32_t state);

/* 1 if the EOF step of the lowered DFA sees a match in `state` (the rc an
 * sre_vm_thompson_exec(ctx, NULL, 0, eof=1) call would add: SRE_OK vs DECLINED) */
SRE_API int sre_cuda_dfa_fin(sre_cuda_program_t *cp, uint32_t state);

/* Host-buffer conveniences (the end-to-end path: H2D + kernels + D2H inside) */
/* sre_cuda_thompson_exec_stream over a stream in host memory (pinned memory for
 * full PCIe speed): slices of slice_bytes (0: 256 MiB; rounded to whole chunks)
 * are copied while the previous slice is scanned; stops at the first match */
SRE_API int sre_cuda_thompson_exec_stream_host(sre_cuda_program_t *cp,
    const uint8_t *host_buf, size_t len, size_t chunk_bytes, unsigned eof,
    uint32_t *state_io, int64_t *match_chunk, size_t slice_bytes);

SRE_API int sre_cuda_thompson_exec_lines_host(sre_cuda_program_t *cp,
    const uint8_t *host_buf, size_t nlines, size_t pitch, size_t linelen,
    int32_t *host_rc, int engine);
SRE_API int sre_cuda_pike_exec_lines_host(sre_cuda_program_t *cp,
    const uint8_t *host_buf, size_t nlines, size_t pitch, size_t linelen,
    int gate_with_thompson, int32_t *host_rc, int64_t *host_ovec,
    size_t ovec_slots);

/*
 * Line index of a '\n'-delimited device buffer, for the ragged entry points
 * above (new: the reference is handed one buffer per exec call by its caller).
 * dev_offsets[0] = 0, dev_offsets[i+1] = one past the '\n' that ends line i; a
 * last line without '\n' ends at len.  So line i = [dev_offsets[i],
 * dev_offsets[i+1]) includes its terminator (`$` still matches before it,
 * sre_vm_thompson.c:183-188).  dev_offsets has room for max_lines + 1 values;
 * *nlines = lines found, which may exceed max_lines (then only the first
 * max_lines are indexed: call again with a larger array).  Synchronises the
 * stream.
 */
SRE_API int sre_cuda_index_lines(const uint8_t *dev_buf, size_t len,
    int64_t *dev_offsets, size_t max_lines, size_t *nlines, void *stream);

/*
 * grep: the Thompson verdict of EVERY line of a '\n'-delimited device buffer in one
 * pass over the text -- sre_vm_thompson_exec(ctx, line, len, eof=1) with a fresh
 * ctx per line, without a line index made beforehand.  dev_rc[i] = SRE_OK or
 * SRE_DECLINED for line i < max_lines; dev_offsets (may be NULL; room for
 * max_lines + 1) receives what sre_cuda_index_lines would: line i =
 * [dev_offsets[i], dev_offsets[i+1]) with its terminator, a last line without
 * '\n' ends at len.  *nlines = lines found (may exceed max_lines: then only the
 * first max_lines rows are written).  Programs whose DFA has a byte table of at
 * most 128 states take the one-pass kernel (dev_buf 16-byte aligned); the others
 * go through sre_cuda_index_lines + sre_cuda_thompson_exec_ragged.  With
 * dev_offsets == NULL and at most 64 states the verdict-only kernel runs (about
 * twice as fast: no per-line records).  Synchronises the stream.
 * Environment (tuning / tests): SRE_CUDA_TEXT_PIECE = bytes of text per CUDA
 * thread in the verdict-only kernel (a multiple of 128 in [512, 4096]; default:
 * fitted to the input).
 */
SRE_API int sre_cuda_thompson_exec_text(sre_cuda_program_t *cp, const uint8_t *dev_buf,
    size_t len, int64_t *dev_offsets, int32_t *dev_rc, size_t max_lines, size_t *nlines,
    void *stream);

/* Tuning / introspection */
/* Pike tier sre_cuda_pike_exec_lines may use for this program (tests): 0 = best
 * available (default): the determinised Pike VM when the program has one, else
 * the closure-table kernel, each followed by the next tier for the lines it
 * gives up on; 1 = general kernel only; 2 = walking shared-memory kernel; 3 =
 * closure-table kernel also when the program has a determinised form */
SRE_API void sre_cuda_program_set_pike_tier(sre_cuda_program_t *cp, int mode);
/* the first tier of the last sre_cuda_pike_exec_lines call on the program: 3 = determinised
 * Pike VM (k_pike_lineage), 0 = closure tables, 2 = walking, 1 = general; -1: none yet */
SRE_API int sre_cuda_program_last_pike_tier(sre_cuda_program_t *cp);
SRE_API long sre_cuda_launch_count(int reset);      /* kernels launched so far   */
SRE_API int sre_cuda_device_available(void);        /* 1 if a CUDA device works  */
SRE_API const char *sre_cuda_last_error(void);      /* of the calling thread     */

#ifdef __cplusplus
}
#endif

#endif /* SREGEX_B200_SREGEX_CUDA_H */
