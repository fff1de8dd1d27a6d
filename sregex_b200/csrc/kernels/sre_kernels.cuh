/*
 * sre_kernels.cuh -- device-side table layouts and kernel launch prototypes
 * of libsregex_cuda (sm_100a).  See DESIGN.md section 4 for each kernel's
 * bound and algorithmic bytes.
 */
#ifndef SRE_KERNELS_CUH
#define SRE_KERNELS_CUH

#include <stdint.h>
#include <cuda_runtime.h>

/* status codes as seen by kernels (== include/sregex/sregex.h) */
#define SRE_K_OK        0
#define SRE_K_ERROR    (-1)
#define SRE_K_AGAIN    (-2)
#define SRE_K_DECLINED (-5)
#define SRE_K_RETRY    (-100)  /* internal: line must be re-run by the general kernel */
#define SRE_K_QUIRK    (-101)  /* internal: the reference's prefilter misfire is possible on this line:
                                  re-run by the general kernel in faithful mode (k_pike_quirk_mark)   */

/* ---- DFA tier ------------------------------------------------------------ */
struct sre_dev_dfa_t {
    uint32_t         nstates, nclasses, start, acc;
    const uint8_t   *t256;      /* [nstates][256] next state (u8), or NULL    */
    const uint16_t  *tcls;      /* [nstates][nclasses] next state * nclasses  */
    const uint8_t   *clsmap;    /* [256]                                      */
    const uint8_t   *fin;       /* [nstates]                                  */
    const uint8_t   *h256;      /* [256][256] next | restart flag, or NULL    */
    const uint8_t   *x256;      /* [256][256] text table (sre_text.cu), or NULL: as t256, but '\n' leads to
                                   start | 0x80 when the line that ends there matched */
    const uint8_t   *x256m;     /* the same for <= 64 states, '\n' also sets bit 6 (rows r + 64k alike), or NULL */
    uint32_t         xguess;    /* the state the automaton idles in on ordinary text: what k_text_verdicts
                                   enters a piece with when a line is open there (a guess, checked later) */
    const uint16_t  *hcls;      /* [nstates][hncls] next | 0x8000 restart, or NULL */
    const uint8_t   *hclsmap;   /* [256]                                      */
    uint32_t         hncls;
};

/* ---- NFA tier ------------------------------------------------------------ */
struct sre_dev_nfa_t {
    uint32_t         nstates, nwords, nclasses, nkinds;
    const uint8_t   *clsmap;    /* [256]                                      */
    const uint8_t   *cls_kind;  /* [nclasses]                                 */
    const uint32_t  *mv;        /* [nclasses][nwords]                         */
    const uint32_t  *mt;        /* [nclasses][nwords]                         */
    const uint32_t  *mt_eof;    /* [nwords]                                   */
    const uint32_t  *init;      /* [nwords]                                   */
    const uint32_t  *shift_mask;/* [nwords]                                   */
    const uint32_t  *follow;    /* [nkinds][nrows][nwords], complex rows only */
    const int32_t   *rowidx;    /* [nstates] -> row or -1                     */
    uint32_t         nrows;
    const uint32_t  *any_follow;   /* [nkinds][nwords] follow row of the ".*?" any state */
    const uint32_t  *complex_mask; /* [nwords] movers that need a follow-row OR          */
    const uint32_t  *match_mask;   /* [nwords] MATCH states                              */
    uint32_t         match_lookahead;  /* some MATCH state has a pending look-ahead     */
};

/* the same for programs of at most 64 lowered states: bitsets as 64-bit words
 * (thread-per-line kernel k_nfa64_lines; nstates == 0: not available) */
struct sre_dev_nfa64_t {
    uint32_t          nstates, nclasses, nkinds;
    const uint8_t    *clsmap;       /* [256]                                   */
    const uint8_t    *cls_kind;     /* [nclasses]                              */
    const uint64_t   *mv, *mt;      /* [nclasses]                              */
    const uint64_t   *follow;       /* [nkinds][64] rows of the non-shift movers */
    uint64_t          init, mt_eof, shift_mask, complex_mask, any_follow[3];
    /* the non-shift movers by number when there are at most 4 of them (the rule for regexes with
     * bounded repetitions: one loop state and a run of shift states): the kernel then ORs their
     * follow rows in under a mask instead of looping over set bits */
    uint32_t          ncomplex;         /* 0xffffffff: more than 4 */
    uint8_t           cidx[4];
};
cudaError_t sre_launch_nfa64_lines(const sre_dev_nfa64_t &nfa, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int32_t *rc, cudaStream_t stream, int *launches);

/* ---- Pike tier (runs the bytecode itself) -------------------------------- */
struct sre_dev_inst_t {         /* == host sre_instruction_t, 16 bytes        */
    uint8_t   opcode, ch;
    uint16_t  nranges;
    int32_t   x, y, v;
};

/* one thread of the start closure (threads that closure(pc 0) parks on a
 * consuming instruction), with the capture slots the path to it SAVEs */
struct sre_dev_start_t {
    int32_t   pc;
    uint8_t   nsl;          /* slots set to the current position ...          */
    uint8_t   sl[7];        /* ... relative to the owning regex's first slot  */
};

struct sre_dev_pike_t {
    uint32_t                 len;           /* instructions                   */
    uint32_t                 nslots;        /* capture slots, all regexes     */
    uint32_t                 nregexes;
    uint32_t                 nleading;
    int32_t                  leading_byte;
    const sre_dev_inst_t    *insts;
    const uint8_t           *ranges;        /* (from,to) pairs                */
    const int32_t           *leading;       /* pcs of leading instructions    */
    const uint32_t          *slot_ofs;      /* [nregexes+1] first slot        */
    const uint16_t          *pc_regex;      /* [len] regex that owns the pc   */
    uint32_t                 max_slots;     /* slots of the largest regex     */
    uint32_t                 max_threads;   /* thread pool entries per ctx    */
    uint32_t                 stack_cap;     /* DFS stack entries per ctx      */
    uint64_t                 ctx_stride;    /* bytes of scratch per ctx       */
    uint32_t                 leadset[8];    /* bytes some leading inst takes  */
    /* the reference's prefilter misfire (lower/sre_quirk.h): possible at all; the byte values a
     * match must start with, right after a prefilter jump, to trigger it */
    uint32_t                 quirk_possible, quirk_single[8];
    /* start closure by next byte: entries [start_ofs[b], start_ofs[b+1]) in
     * priority order; NULL when closure(pc 0) meets an assertion             */
    const uint32_t          *start_ofs;
    const sre_dev_start_t   *start_ent;
    /* closure tables of k_pike_table (lower/sre_closure.h; clo_nent == 0:
     * none), over the clo_npark instructions a thread can be parked on       */
    const uint32_t          *clo_ent;       /* parked number                       */
    const uint32_t          *clo_emask;     /* slots SAVEd on the path to it       */
    const uint16_t          *clo_ofs;       /* [3][npark + 2]                      */
    const uint32_t          *clo_accept;    /* [nsets][8] distinct byte sets       */
    const uint16_t          *clo_accidx;    /* [npark] byte set it takes           */
    const uint16_t          *clo_regex;     /* [npark] owning regex                */
    const uint8_t           *clo_kind;      /* [npark] what it is                  */
    const uint32_t          *clo_bent;      /* start closure bucketed by next byte */
    const uint32_t          *clo_bmask;
    const uint16_t          *clo_bofs;      /* [3][257]                            */
    uint32_t                 clo_nent, clo_nbent, clo_nsets, clo_npark;
    uint32_t                 clo_p_any;     /* parked number of the ".*?" ANY      */
    uint32_t                 clo_has_hold;  /* some parked instruction is a look-ahead assertion */
    uint32_t                 clo_ctx_dep;   /* program has \A or ^             */
};

/* launchers (sre_kernels.cu); all asynchronous on `stream` ------------------ */

/* line i = buf[i*pitch, i*pitch+linelen); needs buf%16==0 && pitch%16==0     */
cudaError_t sre_launch_dfa_lines(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    size_t nlines, size_t pitch, size_t linelen, int32_t *rc, int variant,
    cudaStream_t stream, int *launches);

/* ragged / unaligned lines: line i = buf[off[i], off[i+1]) or fixed pitch    */
cudaError_t sre_launch_dfa_ragged(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen,
    int32_t *rc, cudaStream_t stream, int *launches);

/*
 * Bit-parallel NFA, one warp per line.  state_io (nlines*nwords u32, may be
 * NULL) carries the thread set across calls when from_init == 0 / for
 * SRE_AGAIN; rc is SRE_OK / SRE_DECLINED (eof) / SRE_AGAIN (!eof).
 */
cudaError_t sre_launch_nfa_lines(const sre_dev_nfa_t &nfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen,
    uint32_t *state_io, int from_init, int eof, int32_t *rc,
    cudaStream_t stream, int *launches);

/* serial DFA with state carry (one thread per line; the classic exec path)   */
cudaError_t sre_launch_dfa_carry(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen,
    uint32_t *state_io, int from_init, int eof, int32_t *rc,
    cudaStream_t stream, int *launches);

/* skip-scan flavour: start state left by <= 4 byte values (pats: each value
 * replicated into the 4 bytes of a word)                                       */
cudaError_t sre_launch_dfa_lines_skip(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    size_t nlines, size_t pitch, size_t linelen, int32_t *rc, const uint32_t *pats,
    int npat, int variant, cudaStream_t stream, int *launches);

/* Thompson verdict + Pike start hint per line (needs dfa.h256, aligned lines):
 * hint[i] = offset after which no earlier-started thread is alive; pats / npat
 * (may be NULL / 0): the 1 or 2 byte values that leave the start state, as
 * for sre_launch_dfa_lines_skip, to skip words no lane needs                  */
/* What a gate kernel does for the Pike pass behind it besides verdict and hint (list == NULL:
 * nothing): the lines whose verdict is SRE_OK are appended to list[*count ...] (any order), the
 * ovector rows of the others are set to -1 (ovec may be NULL). */
struct sre_gate_pack_t {
    uint32_t *list, *count;
    int64_t  *ovec;
    uint32_t  ovec_slots;
};
cudaError_t sre_launch_dfa_lines_hint(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    size_t nlines, size_t pitch, size_t linelen, int32_t *rc, int32_t *hint,
    const uint32_t *pats, int npat, sre_gate_pack_t pack, cudaStream_t stream, int *launches);

/* same for any DFA size / alignment / ragged offsets (thread per line, dfa.hcls) */
cudaError_t sre_launch_dfa_generic_hint(const sre_dev_dfa_t &dfa, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, int32_t *rc,
    int32_t *hint, cudaStream_t stream, int *launches);

/* TMA-tiled lines, class table through L1/L2 (tables beyond shared memory); hint may be NULL */
cudaError_t sre_launch_dfa_lines_big(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int32_t *rc, int32_t *hint, sre_gate_pack_t pack, int variant, cudaStream_t stream,
    int *launches);

/* the lines a Pike pass works on: list[0 .. *count) (device memory), or every
 * line 0 .. nlines-1 when list == NULL */
struct sre_line_list_t {
    const uint32_t *list;
    const uint32_t *count;
};

/* list <- lines with select[line] == SRE_K_OK, in no particular order; the
 * other lines get rc[line] = select[line].  *count must be 0 beforehand.       */
cudaError_t sre_launch_pike_compact(const int32_t *select, size_t nlines, int32_t *rc,
    uint32_t *list, uint32_t *count, cudaStream_t stream, int *launches);

/* Pike VM over the listed lines (rows of rc / ovec not listed are not touched);
 * start may be NULL or per-line offsets at which the search may begin          */
size_t sre_pike_concurrency(size_t nlines);
cudaError_t sre_launch_pike_lines(const sre_dev_pike_t &pk, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen,
    sre_line_list_t lines, const int32_t *start, int32_t *rc, int64_t *ovec,
    uint32_t ovec_slots, uint8_t *scratch, size_t nctx, int retry_only, cudaStream_t stream,
    int *launches, const uint32_t *retry_count = nullptr);      /* retry_count: device, 0 = nothing to re-run */

/* shared-memory Pike for small single-regex programs; lines that exceed its
 * capacities get rc = SRE_K_RETRY (re-run them with retry_only = 1 above)      */
bool sre_pike_small_applicable(const sre_dev_pike_t &pk);
cudaError_t sre_launch_pike_small(const sre_dev_pike_t &pk, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen,
    sre_line_list_t lines, const int32_t *start, int32_t *rc, int64_t *ovec,
    uint32_t ovec_slots, cudaStream_t stream, int *launches);

/* device-side bookkeeping of one sre_cuda_pike_exec_lines call: the table
 * kernel's work counter, and how many lines each table pass gave up on (so
 * that the passes after it return at once when there is nothing to re-run)    */
struct sre_pike_work_t {
    unsigned long long next;
    uint32_t           given_up[2];
};

/* closure-table Pike for small single-regex programs (sre_pike_table.cu);
 * same contract as sre_launch_pike_small.  K threads per list, H pending
 * look-ahead closures per context; retry_only: only lines with rc RETRY;
 * pass: 0 first pass, 1 the retry pass; work: device memory, see above       */
bool sre_pike_table_applicable(const sre_dev_pike_t &pk, const int64_t *offsets, size_t linelen,
    int K, int H);             /* for these lines, with lists of K / H */
cudaError_t sre_launch_pike_table(const sre_dev_pike_t &pk, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen,
    sre_line_list_t lines, const int32_t *start, int32_t *rc, int64_t *ovec,
    uint32_t ovec_slots, int K, int H, int pass, sre_pike_work_t *work,
    cudaStream_t stream, int *launches);

/* the determinised Pike VM (lower/sre_pdfa.h; nstates == 0: none) and its capture kernel
 * (sre_pike_lineage.cu): same contract as sre_launch_pike_table, pass 0 */
struct sre_dev_pdfa_t {
    uint32_t         nstates, nclasses, max_slots;
    uint32_t         nent;          /* provenance records                      */
    /* the start closure by what lies in front of the first byte (0: nothing, 1: a newline, 2: a
     * word byte, 3: anything else; all alike unless ctx_dep): its state, the index of the ".*?"
     * thread in it (0xff: none), where its slots begin in init_mask */
    uint32_t         init[4], init_any[4], init_mask_ofs[4], ctx_dep;
    const uint8_t   *clsmap;        /* [256]                                   */
    const uint16_t  *trans;         /* [nstates][nclasses] next | 0x8000 match */
    /* provenance of the threads of a transition's next list: record eofs[t] + j =
     * { slots SAVEd (they take the position after the consumed byte),
     *   parent index | 0x100 when the parent is the ".*?" thread (the lineage ends) } */
    const uint32_t  *eofs;          /* [nstates * nclasses + 1]                */
    const uint2     *ent;           /* [nent]                                  */
    /* the thread that matched in a transition with bit 15: { slots, parent | 0x100 | regex << 16 } */
    const uint2     *mev;           /* [nstates * nclasses]                    */
    const uint32_t  *eof;           /* [nstates] first parked MATCH thread (EOF step): index | regex << 16, 0xff: none */
    const uint32_t  *init_mask;     /* slots SAVEd by a start closure, per thread */
    /* programs with look-ahead assertions (NULL otherwise): the slots SAVEd at the position of
     * the step itself, by assertions resolved on the way (lower/sre_pdfa.h: emask0 / mmask0 /
     * eof_mask0), next to ent / mev / eof */
    const uint32_t  *ent0, *mev0, *eof0;
};
bool sre_pike_lineage_applicable(const sre_dev_pdfa_t &d, size_t linelen);
cudaError_t sre_launch_pike_lineage(const sre_dev_pdfa_t &d, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines, const int32_t *start, int32_t *rc,
    int64_t *ovec, uint32_t ovec_slots, int prefilled, sre_pike_work_t *work, cudaStream_t stream, int *launches);

/* Lines on which the reference's first-byte prefilter can misfire (sre_quirk.h) get rc =
 * SRE_K_QUIRK and are counted in *count: matched lines whose match starts at offset s >= 1 with a
 * quirk_single byte, a non-leading byte in front of it and a non-leading byte behind it.  With
 * ovec == NULL (or no slots) every matched line is marked.  Re-run them with
 * sre_launch_pike_lines(retry_only = 2): the reference to the letter, from offset 0. */
cudaError_t sre_launch_pike_quirk_mark(const sre_dev_pike_t &pk, const uint8_t *buf, const int64_t *offsets,
    size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines, int32_t *rc, const int64_t *ovec,
    uint32_t ovec_slots, uint32_t *count, cudaStream_t stream, int *launches);

/* all non-overlapping matches per line (post-match continuation, global scan)  */
cudaError_t sre_launch_pike_lines_all(const sre_dev_pike_t &pk, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, uint32_t max_matches,
    int32_t *count, int64_t *spans, int32_t *ids, uint8_t *scratch, size_t nctx,
    cudaStream_t stream, int *launches);

/* line index of a '\n'-delimited buffer (sre_lines.cu): offsets[0] = 0,
 * offsets[i+1] = one past the newline ending line i (at most max_lines of them);
 * a last line without newline ends at len; workspace
 * (sre_lines_workspace_bytes(len) bytes): its last word receives the number of
 * lines found (which may exceed max_lines)                                      */
size_t sre_lines_workspace_bytes(size_t len);
cudaError_t sre_launch_index_lines(const uint8_t *buf, size_t len, int64_t *offsets, size_t max_lines,
    unsigned long long *workspace, cudaStream_t stream, int *launches);

/* every line of a '\n'-delimited buffer in one pass (sre_text.cu; needs dfa.x256, buf 16-byte
 * aligned): rc[line], offsets[line + 1] (offsets may be NULL) for line < max_lines; the number
 * of lines is left at workspace + sre_text_count_offset(len) (8 bytes) */
size_t sre_text_workspace_bytes(size_t len);
size_t sre_text_count_offset(size_t len);
cudaError_t sre_launch_text(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len, int32_t *rc,
    int64_t *offsets, size_t max_lines, uint8_t *workspace, cudaStream_t stream, int *launches);

/* Pike VM streaming step on one persistent context (classic API)             */
/* skip: with a fresh context, begin at buf + skip as if the bytes before had been fed in an
 * earlier call and left no thread alive (offsets stay relative to buf) */
cudaError_t sre_launch_pike_stream(const sre_dev_pike_t &pk, uint8_t *ctx,
    const uint8_t *buf, size_t len, size_t skip, int eof, int want_pending, int64_t *out,
    uint32_t ovec_slots, cudaStream_t stream, int *launches);
cudaError_t sre_launch_pike_ctx_init(const sre_dev_pike_t &pk, uint8_t *ctx,
    cudaStream_t stream, int *launches);
/* many persistent contexts at once (ctxs: nstreams * pk.ctx_stride bytes): stream i is fed
 * buf[offsets[i], offsets[i+1]); out row i = { rc, pending flag, pending[2], ovector[ovec_slots] } */
cudaError_t sre_launch_pike_streams_init(const sre_dev_pike_t &pk, uint8_t *ctxs, size_t nstreams,
    cudaStream_t stream, int *launches);
cudaError_t sre_launch_pike_streams(const sre_dev_pike_t &pk, uint8_t *ctxs, size_t nstreams, const uint8_t *buf,
    const int64_t *offsets, const uint8_t *eofs, int eof_all, int64_t *out, uint32_t ovec_slots,
    cudaStream_t stream, int *launches);

/* chunk-parallel DFA stream scan (sre_stream.cu) ----------------------------- */

#define SRE_STREAM_K         8              /* candidate entry states per record        */
#define SRE_STREAM_FN_BYTES  32             /* one record: K candidates + K exit states */
#define SRE_STREAM_HALO      256            /* bytes of the preceding part a part needs */
#define SRE_STREAM_UNKNOWN   0xffffffffu    /* entry state not known                    */

/* image automaton of the DFA (lower/sre_image.h), over the DFA's byte classes */
struct sre_dev_image_t {
    uint32_t         nstates, nclasses, K;
    const uint16_t  *trans;     /* [nstates][nclasses]                        */
    const uint16_t  *cand;      /* [nstates][K], 16-byte aligned rows         */
    const uint8_t   *ncand;     /* [nstates] 0..K, 0xff: wide                 */
};

struct sre_stream_ws_t {        /* workspace owned by the caller              */
    uint8_t  *fn[4];            /* records per level                          */
    size_t    count[4];
    unsigned long long *first_acc;  /* first piece in which ACC is entered    */
};
size_t sre_stream_piece_bytes(void);
uint32_t sre_stream_fan(void);
cudaError_t sre_launch_dfa_stream_reduce(const sre_dev_dfa_t &dfa, const sre_dev_image_t &img,
    const uint8_t *buf, size_t len, const uint8_t *halo, uint32_t entry, const sre_stream_ws_t &ws,
    cudaStream_t stream, int *launches);
cudaError_t sre_launch_dfa_stream_fix(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
    uint32_t entry_state, const sre_stream_ws_t &ws, unsigned long long *dev_fixed, cudaStream_t stream,
    int *launches);
cudaError_t sre_launch_dfa_stream_walk(const sre_dev_dfa_t &dfa, const sre_dev_image_t &img,
    const uint8_t *buf, size_t len, uint32_t entry_state, const sre_stream_ws_t &ws, uint32_t *dev_out,
    long long *dev_match_offset, cudaStream_t stream, int *launches);
cudaError_t sre_launch_dfa_stream_restart(const sre_dev_dfa_t &dfa, const uint8_t *buf, size_t len,
    uint32_t entry_state, size_t limit, const sre_stream_ws_t &ws, long long *dev_out, cudaStream_t stream,
    int *launches);
/* host side of the record format */
uint32_t sre_stream_fn_apply(const uint8_t *fn, uint32_t state);    /* SRE_STREAM_UNKNOWN: not known */
void sre_stream_fn_compose(uint8_t *fn, const uint8_t *then);       /* fn = then o fn                */
void sre_stream_fn_identity(uint8_t *fn);
int sre_stream_fn_unresolved(const uint8_t *fn);
size_t sre_pike_ctx_bytes(uint32_t len, uint32_t nslots, uint32_t max_slots, uint32_t nthreads,
    uint32_t stack_cap);

#endif
