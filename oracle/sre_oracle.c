/*
 * sre_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the
 * reference's two match executors, used as the parity checker for the CUDA
 * path.  Nothing in the product (libsregex_cuda) may link, call or fall back
 * to this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 * tests/golden/t_suite.json.gz, i.e. the outputs of the unmodified reference
 * (oracle/_ref, built from /root/reference) on all runnable blocks of the
 * reference's t/ suite -- Thompson rc, Pike rc + full ovector, and the
 * 1-byte-chunk streaming sequences incl. temp captures / pending matches --
 * and, when oracle/_ref is present, live against the reference on seeded
 * random inputs.
 *
 * It follows the reference's *algorithm* function by function (citations at
 * each function) on our own pointer-free program representation
 * (sre_internal.h); run-time dedup tags live in the context instead of inside
 * the program (the reference writes pc->tag, sre_vm_thompson.c:284), and Pike
 * captures are plain per-thread value copies instead of ref-counted
 * copy-on-write vectors (sre_capture.c:59-85) -- the observable results are the
 * same, which is what the golden vectors pin.
 *
 * Exports the sregex.h VM entry points so the same ctypes harness drives the
 * reference, this oracle and libsregex_cuda.
 */
#define _GNU_SOURCE
#include "sre_internal.h"
#include <pthread.h>
#include <time.h>

/* ========================================================================
 * Thompson VM (sre_vm_thompson.c)
 * ====================================================================== */

typedef struct {
    int32_t   pc;
    uint8_t   seen_word;
} ot_thread_t;

typedef struct {
    uint32_t      count, cap;
    ot_thread_t  *threads;
} ot_list_t;

struct sre_vm_thompson_ctx_s {
    sre_pool_t      *pool;
    sre_program_t   *prog;
    const sre_char  *buffer;
    ot_list_t        lists[2];
    ot_list_t       *clist, *nlist;
    unsigned         tag;
    unsigned        *tags;      /* per pc; the reference keeps these in prog */
    uint8_t          first_buf;
};

static int
ot_list_init(sre_pool_t *pool, ot_list_t *l, uint32_t cap)
{
    l->count = 0;
    l->cap = cap;
    l->threads = sre_palloc(pool, cap * sizeof(ot_thread_t));
    return l->threads ? SRE_OK : SRE_ERROR;
}

/* sre_vm_thompson_create_ctx, sre_vm_thompson.c:25-60 */
SRE_API sre_vm_thompson_ctx_t *
sre_vm_thompson_create_ctx(sre_pool_t *pool, sre_program_t *prog)
{
    sre_vm_thompson_ctx_t *ctx = sre_pcalloc(pool, sizeof(*ctx));
    if (ctx == NULL) {
        return NULL;
    }
    ctx->pool = pool;
    ctx->prog = prog;
    /* the reference sizes both lists to prog->len; the tag-- trick of
     * assertion_hold can re-add a pc, so we leave headroom */
    if (ot_list_init(pool, &ctx->lists[0], 4 * prog->len + 8) != SRE_OK
        || ot_list_init(pool, &ctx->lists[1], 4 * prog->len + 8) != SRE_OK)
    {
        return NULL;
    }
    ctx->clist = &ctx->lists[0];
    ctx->nlist = &ctx->lists[1];
    ctx->tags = sre_pcalloc(pool, (prog->len + 1) * sizeof(unsigned));
    if (ctx->tags == NULL) {
        return NULL;
    }
    ctx->tag = 1;
    ctx->first_buf = 1;
    return ctx;
}

static int
oracle_in_ranges(sre_program_t *prog, sre_instruction_t *in, sre_char c)
{
    uint32_t j;
    for (j = 0; j < in->nranges; j++) {
        if (c >= prog->ranges[in->v + j].from && c <= prog->ranges[in->v + j].to) {
            return 1;
        }
    }
    return 0;
}

/* sre_vm_thompson_add_thread, sre_vm_thompson.c:273-345 */
static void
ot_add_thread(sre_vm_thompson_ctx_t *ctx, ot_list_t *l, int32_t pc,
    const sre_char *sp)
{
    sre_instruction_t  *in = &ctx->prog->insts[pc];
    uint8_t             seen_word = 0;
    ot_thread_t        *t;

    if (ctx->tags[pc] == ctx->tag) {
        return;                             /* already on list, :281 */
    }
    ctx->tags[pc] = ctx->tag;

    switch (in->opcode) {
    case SRE_OPCODE_JMP:
        ot_add_thread(ctx, l, in->x, sp);
        return;
    case SRE_OPCODE_SPLIT:
        ot_add_thread(ctx, l, in->x, sp);
        ot_add_thread(ctx, l, in->y, sp);
        return;
    case SRE_OPCODE_SAVE:
        ot_add_thread(ctx, l, pc + 1, sp);
        return;
    case SRE_OPCODE_ASSERT:
        switch (in->v) {
        case SRE_REGEX_ASSERT_BIG_A:        /* :301-308 */
            if (sp != ctx->buffer) {
                return;
            }
            ot_add_thread(ctx, l, pc + 1, sp);
            return;
        case SRE_REGEX_ASSERT_CARET:        /* :310-316 */
            if (sp != ctx->buffer && sp[-1] != '\n') {
                return;
            }
            ot_add_thread(ctx, l, pc + 1, sp);
            return;
        case SRE_REGEX_ASSERT_SMALL_B:
        case SRE_REGEX_ASSERT_BIG_B:        /* :318-325 */
            seen_word = (sp != ctx->buffer && sre_isword(sp[-1]));
            break;
        default:                            /* look-ahead: postponed */
            break;
        }
        break;
    default:
        break;
    }

    if (l->count >= l->cap) {
        return;     /* cannot happen with the headroom above */
    }
    t = &l->threads[l->count++];
    t->pc = pc;
    t->seen_word = seen_word;
}

/* sre_vm_thompson_exec, sre_vm_thompson.c:63-270 */
SRE_API sre_int_t
sre_vm_thompson_exec(sre_vm_thompson_ctx_t *ctx, sre_char *input, size_t size,
    unsigned eof)
{
    sre_program_t      *prog = ctx->prog;
    ot_list_t          *clist = ctx->clist, *nlist = ctx->nlist, *tmp;
    const sre_char     *sp, *last;
    sre_instruction_t  *in;
    ot_thread_t        *t;
    uint32_t            i;
    int                 cur_word, hold;

    ctx->buffer = input;
    if (ctx->first_buf) {
        ctx->first_buf = 0;
        ot_add_thread(ctx, clist, 0, input);
    }
    last = input + size;

    for (sp = input; sp < last || (eof && sp == last); sp++) {
        if (clist->count == 0) {
            break;
        }
        ctx->tag++;

        for (i = 0; i < clist->count; i++) {    /* list may grow, :100 */
            t = &clist->threads[i];
            in = &prog->insts[t->pc];

            switch (in->opcode) {
            case SRE_OPCODE_IN:
            case SRE_OPCODE_NOTIN:
                if (sp == last) {
                    break;
                }
                if (oracle_in_ranges(prog, in, *sp)
                    != (in->opcode == SRE_OPCODE_IN))
                {
                    break;
                }
                ot_add_thread(ctx, nlist, t->pc + 1, sp + 1);
                break;
            case SRE_OPCODE_CHAR:
                if (sp == last || *sp != in->ch) {
                    break;
                }
                ot_add_thread(ctx, nlist, t->pc + 1, sp + 1);
                break;
            case SRE_OPCODE_ANY:
                if (sp == last) {
                    break;
                }
                ot_add_thread(ctx, nlist, t->pc + 1, sp + 1);
                break;
            case SRE_OPCODE_ASSERT:             /* :174-231 */
                cur_word = (sp != last && sre_isword(*sp));
                hold = 0;
                switch (in->v) {
                case SRE_REGEX_ASSERT_SMALL_Z:
                    hold = (sp == last);
                    break;
                case SRE_REGEX_ASSERT_DOLLAR:
                    hold = (sp == last || *sp == '\n');
                    break;
                case SRE_REGEX_ASSERT_BIG_B:
                    hold = !(t->seen_word ^ cur_word);
                    break;
                case SRE_REGEX_ASSERT_SMALL_B:
                    hold = (t->seen_word ^ cur_word);
                    break;
                default:
                    break;
                }
                if (hold) {
                    int32_t pc = t->pc;     /* t may move? no: array fixed */
                    ctx->tag--;
                    ot_add_thread(ctx, clist, pc + 1, sp);
                    ctx->tag++;
                }
                break;
            case SRE_OPCODE_MATCH:
                return SRE_OK;
            default:
                break;
            }
        }

        tmp = clist;
        clist = nlist;
        nlist = tmp;
        nlist->count = 0;
        if (sp == last) {
            break;
        }
    }

    ctx->clist = clist;
    ctx->nlist = nlist;
    return eof ? SRE_DECLINED : SRE_AGAIN;
}

/* ========================================================================
 * Pike VM (sre_vm_pike.c)
 * ====================================================================== */

typedef struct op_thread_s  op_thread_t;
struct op_thread_s {
    int32_t        pc;
    unsigned       seen_word;
    op_thread_t   *next;
    sre_int_t      regex_id;    /* lives in sre_capture_t in the reference  */
    sre_int_t      cap[1];      /* nslots values                            */
};

typedef struct {
    uint32_t       count;
    op_thread_t   *head, **tailp;
} op_list_t;

struct sre_vm_pike_ctx_s {
    unsigned         tag, prog_tag;
    unsigned        *tags;
    sre_int_t        processed_bytes;
    const sre_char  *buffer;
    sre_pool_t      *pool;
    sre_program_t   *prog;
    uint32_t         nslots;
    int              has_matched;
    sre_int_t        matched_id;
    sre_int_t       *matched;           /* nslots                            */
    sre_int_t       *done_cap;          /* capture handed out by SRE_DONE    */
    sre_int_t        done_id;
    op_thread_t     *free_threads;
    sre_int_t        pending_ovector[2];
    sre_int_t       *ovector;
    size_t           ovecsize;
    op_list_t        lists[2], *clist, *nlist;
    sre_int_t        last_matched_pos;
    int32_t         *initial_states;
    uint32_t         initial_states_count;
    unsigned         first_buf, seen_start_state, eof, empty_capture,
                     seen_newline, seen_word;
};

/* test statistic: longest Pike thread list seen (sizing of the GPU kernels) */
static uint32_t oracle_max_list;
/*
 * The reference's first-byte prefilter (sre_vm_pike.c:256-309) is not result
 * neutral: its "is the list the initial one?" test compares the thread COUNT and
 * every pc but the LAST (:262-274, `t && t->next`).  Before a match the last
 * thread is always the ".*?" one, so the test is exact; but once a match has cut
 * that thread, a list of survivors that happens to have the initial list's
 * length and leading pcs -- /a+b?/ after matching one 'a': [a-loop, b] against
 * the initial [a-loop, .*?] -- passes for the initial list.  If that happens
 * right after a prefilter jump (seen_start_state still set), the survivors are
 * dropped and the search starts over at the next leading byte, a later match
 * overwriting the leftmost one: /a+b?/ on ".a.a" reports (3, 4), on " ab" it
 * reports (1, 2).  This oracle restates that faithfully (prefilter on, the
 * default: equal to the reference on every input).  With the prefilter left out
 * it computes what the Pike VM computes without the shortcut -- the leftmost-
 * first match -- which is what libsregex_cuda implements.
 */
static int oracle_prefilter = 1;

SRE_API void oracle_pike_prefilter(int on)
{
    oracle_prefilter = on;
}

SRE_API uint32_t oracle_pike_max_list(int reset)
{
    uint32_t v = oracle_max_list;
    if (reset) {
        oracle_max_list = 0;
    }
    return v;
}

static void
op_list_reset(op_list_t *l)
{
    l->count = 0;
    l->head = NULL;
    l->tailp = &l->head;
}

/* sre_vm_pike_create_ctx, sre_vm_pike.c:94-145 */
SRE_API sre_vm_pike_ctx_t *
sre_vm_pike_create_ctx(sre_pool_t *pool, sre_program_t *prog,
    sre_int_t *ovector, size_t ovecsize)
{
    sre_vm_pike_ctx_t *ctx = sre_pcalloc(pool, sizeof(*ctx));
    if (ctx == NULL) {
        return NULL;
    }
    ctx->pool = pool;
    ctx->prog = prog;
    ctx->nslots = (uint32_t) (prog->ovecsize / sizeof(sre_int_t));
    ctx->tags = sre_pcalloc(pool, (prog->len + 1) * sizeof(unsigned));
    ctx->matched = sre_palloc(pool, prog->ovecsize);
    ctx->done_cap = sre_palloc(pool, prog->ovecsize);
    if (!ctx->tags || !ctx->matched || !ctx->done_cap) {
        return NULL;
    }
    op_list_reset(&ctx->lists[0]);
    op_list_reset(&ctx->lists[1]);
    ctx->clist = &ctx->lists[0];
    ctx->nlist = &ctx->lists[1];
    ctx->ovector = ovector;
    ctx->ovecsize = ovecsize;
    ctx->first_buf = 1;
    return ctx;
}

static void
op_free_thread(sre_vm_pike_ctx_t *ctx, op_thread_t *t)
{
    t->next = ctx->free_threads;
    ctx->free_threads = t;
}

/* sre_vm_pike_clear_thread_list, sre_vm_pike.c:1064-1080 */
static void
op_clear_list(sre_vm_pike_ctx_t *ctx, op_list_t *l)
{
    op_thread_t *t;
    while (l->head) {
        t = l->head;
        l->head = t->next;
        op_free_thread(ctx, t);
    }
    op_list_reset(l);
}

/*
 * sre_vm_pike_add_thread, sre_vm_pike.c:756-942.  `cap` is the caller's
 * private copy and may be modified (value semantics replace the reference's
 * copy-on-write).  want_done mirrors pcap != NULL.
 */
static sre_int_t
op_add_thread(sre_vm_pike_ctx_t *ctx, op_list_t *l, int32_t pc, sre_int_t *cap,
    sre_int_t pos, int want_done)
{
    sre_program_t      *prog = ctx->prog;
    sre_instruction_t  *in = &prog->insts[pc];
    unsigned            seen_word = 0;
    op_thread_t        *t;
    sre_int_t           rc;

    if (ctx->tags[pc] == ctx->tag) {
        /* the revisited-SPLIT rule, :770-786 */
        if (in->opcode == SRE_OPCODE_SPLIT && ctx->tags[in->y] != ctx->tag) {
            if (pc == 0) {
                ctx->seen_start_state = 1;
            }
            return op_add_thread(ctx, l, in->y, cap, pos, want_done);
        }
        return SRE_OK;
    }
    ctx->tags[pc] = ctx->tag;

    switch (in->opcode) {
    case SRE_OPCODE_JMP:
        return op_add_thread(ctx, l, in->x, cap, pos, want_done);

    case SRE_OPCODE_SPLIT: {
        sre_int_t copy[ctx->nslots];
        if (pc == 0) {
            ctx->seen_start_state = 1;
        }
        memcpy(copy, cap, sizeof(copy));
        rc = op_add_thread(ctx, l, in->x, copy, pos, want_done);
        if (rc != SRE_OK) {
            return rc;
        }
        return op_add_thread(ctx, l, in->y, cap, pos, want_done);
    }

    case SRE_OPCODE_SAVE:                   /* :818-837 */
        cap[in->v] = ctx->processed_bytes + pos;
        return op_add_thread(ctx, l, pc + 1, cap, pos, want_done);

    case SRE_OPCODE_ASSERT:
        switch (in->v) {
        case SRE_REGEX_ASSERT_BIG_A:        /* :841-846 */
            if (pos || ctx->processed_bytes) {
                return SRE_OK;
            }
            return op_add_thread(ctx, l, pc + 1, cap, pos, want_done);
        case SRE_REGEX_ASSERT_CARET:        /* :848-864 */
            if (pos == 0) {
                if (ctx->processed_bytes && !ctx->seen_newline) {
                    return SRE_OK;
                }
            } else if (ctx->buffer[pos - 1] != '\n') {
                return SRE_OK;
            }
            return op_add_thread(ctx, l, pc + 1, cap, pos, want_done);
        case SRE_REGEX_ASSERT_SMALL_B:
        case SRE_REGEX_ASSERT_BIG_B:        /* :866-881 */
            seen_word = pos == 0 ? 0 : sre_isword(ctx->buffer[pos - 1]);
            break;
        default:
            break;
        }
        break;

    case SRE_OPCODE_MATCH:                  /* :889-899 */
        ctx->last_matched_pos = cap[1];
        if (want_done) {
            memcpy(ctx->done_cap, cap, prog->ovecsize);
            ctx->done_id = in->v;
            return SRE_DONE;
        }
        break;

    default:
        break;
    }

    if (ctx->free_threads) {
        t = ctx->free_threads;
        ctx->free_threads = t->next;
    } else {
        t = sre_palloc(ctx->pool, sizeof(op_thread_t)
                                  + ctx->nslots * sizeof(sre_int_t));
        if (t == NULL) {
            return SRE_ERROR;
        }
    }
    t->pc = pc;
    t->seen_word = seen_word;
    t->next = NULL;
    t->regex_id = in->opcode == SRE_OPCODE_MATCH ? in->v : 0;
    memcpy(t->cap, cap, prog->ovecsize);
    *l->tailp = t;
    l->tailp = &t->next;
    l->count++;
    if (l->count > oracle_max_list) {
        oracle_max_list = l->count;
    }
    return SRE_OK;
}

/* sre_vm_pike_find_first_byte, sre_vm_pike.c:992-1061 */
static const sre_char *
op_find_first_byte(sre_program_t *prog, const sre_char *pos,
    const sre_char *last)
{
    sre_instruction_t  *in;
    uint32_t            i;

    if (prog->leading_byte != -1) {
        pos = memchr(pos, prog->leading_byte, last - pos);
        return pos ? pos : last;
    }
    for (; pos != last; pos++) {
        for (i = 0; i < prog->nleading; i++) {
            in = &prog->insts[prog->leading[i]];
            if (in->opcode == SRE_OPCODE_CHAR) {
                if (*pos == in->ch) {
                    return pos;
                }
            } else if (oracle_in_ranges(prog, in, *pos)
                       == (in->opcode == SRE_OPCODE_IN))
            {
                return pos;
            }
        }
    }
    return pos;
}

/* sre_vm_pike_prepare_matched_captures, sre_vm_pike.c:945-989 */
static sre_int_t
op_prepare_matched(sre_vm_pike_ctx_t *ctx, sre_int_t *ovector, int complete)
{
    sre_program_t  *prog = ctx->prog;
    sre_uint_t      i, ofs = 0;
    size_t          len;

    if (ctx->matched_id < 0 || (sre_uint_t) ctx->matched_id >= prog->nregexes) {
        return SRE_ERROR;
    }
    for (i = 0; i < (sre_uint_t) ctx->matched_id; i++) {
        ofs += prog->multi_ncaps[i] + 1;
    }
    ofs *= 2;
    len = complete ? 2 * (prog->multi_ncaps[i] + 1) * sizeof(sre_int_t)
                   : 2 * sizeof(sre_int_t);
    memcpy(ovector, &ctx->matched[ofs], len);
    if (complete && ctx->ovecsize > len) {
        memset((char *) ovector + len, -1, ctx->ovecsize - len);
    }
    return SRE_OK;
}

/* sre_vm_pike_prepare_temp_captures, sre_vm_pike.c:692-735 (incl. its use of
 * cap->vector[j + 1] without the per-regex offset, :721) */
static void
op_prepare_temp(sre_vm_pike_ctx_t *ctx)
{
    sre_program_t  *prog = ctx->prog;
    op_thread_t    *t;
    sre_uint_t      i, ofs;
    sre_int_t       a, b;

    ctx->ovector[0] = -1;
    ctx->ovector[1] = -1;
    for (t = ctx->clist->head; t; t = t->next) {
        ofs = 0;
        for (i = 0; i < prog->nregexes; i++) {
            a = ctx->ovector[0];
            b = t->cap[ofs];
            if (b != -1 && (a == -1 || b < a)) {
                ctx->ovector[0] = b;
            }
            a = ctx->ovector[1];
            b = t->cap[1];
            if (b != -1 && (a == -1 || b > a)) {
                ctx->ovector[1] = b;
            }
            ofs += 2 * (prog->multi_ncaps[i] + 1);
        }
    }
}

/* sre_vm_pike_exec, sre_vm_pike.c:148-689 */
SRE_API sre_int_t
sre_vm_pike_exec(sre_vm_pike_ctx_t *ctx, sre_char *input, size_t size,
    unsigned eof, sre_int_t **pending_matched)
{
    sre_program_t      *prog = ctx->prog;
    op_list_t          *clist = ctx->clist, *nlist = ctx->nlist, *tmp, list;
    const sre_char     *sp, *last, *p;
    sre_instruction_t  *in;
    op_thread_t        *t;
    sre_int_t           rc, cap[ctx->nslots];
    unsigned            seen_word, cur_word;
    uint32_t            i;
    int                 hold, take;

    if (ctx->eof) {
        return SRE_ERROR;
    }
    ctx->buffer = input;
    ctx->last_matched_pos = -1;

    if (ctx->empty_capture) {               /* :179-193 */
        ctx->empty_capture = 0;
        if (size == 0) {
            if (eof) {
                ctx->eof = 1;
                return SRE_DECLINED;
            }
            return SRE_AGAIN;
        }
        sp = input + 1;
    } else {
        sp = input;
    }
    last = input + size;

    if (ctx->first_buf) {                   /* :202-229 */
        ctx->first_buf = 0;
        for (i = 0; i < ctx->nslots; i++) {
            cap[i] = -1;
        }
        ctx->tag = ctx->prog_tag + 1;
        rc = op_add_thread(ctx, clist, 0, cap, (sre_int_t) (sp - input), 0);
        if (rc != SRE_OK) {
            ctx->prog_tag = ctx->tag;
            return SRE_ERROR;
        }
        ctx->initial_states_count = clist->count;
        ctx->initial_states = sre_palloc(ctx->pool,
                                         sizeof(int32_t) * (clist->count + 1));
        if (ctx->initial_states == NULL) {
            return SRE_ERROR;
        }
        for (i = 0, t = clist->head; t && t->next; i++, t = t->next) {
            ctx->initial_states[i] = t->pc;
        }
    } else {
        ctx->tag = ctx->prog_tag;
    }

    for (; sp < last || (eof && sp == last); sp++) {
        if (clist->head == NULL) {
            break;
        }

        /* first-byte prefilter, :256-309 (oracle_pike_prefilter(0) leaves it out: see there) */
        if (oracle_prefilter && prog->nleading && ctx->seen_start_state) {
            ctx->seen_start_state = 0;
            if (sp == last || clist->count != ctx->initial_states_count) {
                goto run_cur_threads;
            }
            for (i = 0, t = clist->head; t && t->next; i++, t = t->next) {
                if (t->pc != ctx->initial_states[i]) {
                    goto run_cur_threads;
                }
            }
            p = op_find_first_byte(prog, sp, last);
            if (p > sp) {
                sp = p;
                op_clear_list(ctx, clist);
                for (i = 0; i < ctx->nslots; i++) {
                    cap[i] = -1;
                }
                ctx->tag++;
                rc = op_add_thread(ctx, clist, 0, cap,
                                   (sre_int_t) (sp - input), 0);
                if (rc != SRE_OK) {
                    ctx->prog_tag = ctx->tag;
                    return SRE_ERROR;
                }
                if (sp == last) {
                    break;
                }
            }
        }

run_cur_threads:
        ctx->tag++;

        while (clist->head) {               /* :314-567 */
            t = clist->head;
            clist->head = t->next;
            if (clist->head == NULL) {
                clist->tailp = &clist->head;
            }
            clist->count--;
            in = &prog->insts[t->pc];

            switch (in->opcode) {
            case SRE_OPCODE_IN:
            case SRE_OPCODE_NOTIN:
            case SRE_OPCODE_CHAR:
            case SRE_OPCODE_ANY:
                if (sp == last) {
                    break;
                }
                take = in->opcode == SRE_OPCODE_ANY ? 1
                     : in->opcode == SRE_OPCODE_CHAR ? (*sp == in->ch)
                     : (oracle_in_ranges(prog, in, *sp)
                        == (in->opcode == SRE_OPCODE_IN));
                if (!take) {
                    break;
                }
                rc = op_add_thread(ctx, nlist, t->pc + 1, t->cap,
                                   (sre_int_t) (sp - input + 1), 1);
                if (rc == SRE_DONE) {
                    memcpy(ctx->matched, ctx->done_cap, prog->ovecsize);
                    ctx->matched_id = ctx->done_id;
                    goto matched;
                }
                if (rc != SRE_OK) {
                    ctx->prog_tag = ctx->tag;
                    return SRE_ERROR;
                }
                break;

            case SRE_OPCODE_ASSERT:         /* :449-528 */
                cur_word = (sp != last && sre_isword(*sp));
                seen_word = (t->seen_word || (sp == input && ctx->seen_word));
                hold = 0;
                switch (in->v) {
                case SRE_REGEX_ASSERT_SMALL_Z:
                    hold = (sp == last);
                    break;
                case SRE_REGEX_ASSERT_DOLLAR:
                    hold = (sp == last || *sp == '\n');
                    break;
                case SRE_REGEX_ASSERT_BIG_B:
                    hold = !(seen_word ^ cur_word);
                    break;
                case SRE_REGEX_ASSERT_SMALL_B:
                    hold = (seen_word ^ cur_word);
                    break;
                default:
                    break;
                }
                if (!hold) {
                    break;
                }
                ctx->tag--;
                op_list_reset(&list);
                rc = op_add_thread(ctx, &list, t->pc + 1, t->cap,
                                   (sre_int_t) (sp - input), 0);
                if (rc != SRE_OK) {
                    ctx->prog_tag = ctx->tag + 1;
                    return SRE_ERROR;
                }
                if (list.head) {            /* prepended, :519-523 */
                    if (clist->head == NULL) {
                        clist->tailp = list.tailp;
                    }
                    *list.tailp = clist->head;
                    clist->head = list.head;
                    clist->count += list.count;
                }
                ctx->tag++;
                break;

            case SRE_OPCODE_MATCH:          /* :530-553 */
                ctx->last_matched_pos = t->cap[1];
                memcpy(ctx->matched, t->cap, prog->ovecsize);
                ctx->matched_id = in->v;
matched:
                ctx->has_matched = 1;
                op_free_thread(ctx, t);
                op_clear_list(ctx, clist);
                goto step_done;

            default:
                break;
            }
            op_free_thread(ctx, t);
        }

step_done:
        tmp = clist;
        clist = nlist;
        nlist = tmp;
        if (nlist->head) {
            op_clear_list(ctx, nlist);
        }
        op_list_reset(nlist);
        if (sp == last) {
            break;
        }
    }

    if (ctx->last_matched_pos >= 0) {       /* :586-601 */
        p = input + ctx->last_matched_pos - ctx->processed_bytes;
        if (p > input) {
            ctx->seen_newline = (p[-1] == '\n');
            ctx->seen_word = sre_isword(p[-1]);
        }
        ctx->last_matched_pos = -1;
    }

    ctx->prog_tag = ctx->tag;
    ctx->clist = clist;
    ctx->nlist = nlist;

    if (ctx->has_matched) {
        if (eof || clist->head == NULL) {   /* :607-636 */
            if (op_prepare_matched(ctx, ctx->ovector, 1) != SRE_OK) {
                return SRE_ERROR;
            }
            if (clist->head) {
                op_clear_list(ctx, clist);
                ctx->eof = 1;
            }
            ctx->processed_bytes = ctx->ovector[1];
            ctx->empty_capture = (ctx->ovector[0] == ctx->ovector[1]);
            ctx->has_matched = 0;
            ctx->first_buf = 1;
            return ctx->matched_id;
        }
        if (pending_matched) {              /* :640-658 */
            *pending_matched = ctx->pending_ovector;
            if (op_prepare_matched(ctx, ctx->pending_ovector, 0) != SRE_OK) {
                return SRE_ERROR;
            }
        }
    } else {
        if (eof) {
            ctx->eof = 1;
            return SRE_DECLINED;
        }
        if (pending_matched) {
            *pending_matched = NULL;
        }
    }

    ctx->processed_bytes += (sre_int_t) (sp - input);
    op_prepare_temp(ctx);
    return SRE_AGAIN;
}

/* ========================================================================
 * "JIT" entry points: the oracle has no JIT; like the reference on a
 * non-x86-64 target (sre_vm_thompson_jit.c:43-44) it declines.
 * ====================================================================== */

SRE_API sre_int_t
sre_vm_thompson_jit_compile(sre_pool_t *pool, sre_program_t *prog,
    sre_vm_thompson_code_t **pcode)
{
    (void) pool; (void) prog;
    *pcode = NULL;
    return SRE_DECLINED;
}

SRE_API sre_vm_thompson_ctx_t *
sre_vm_thompson_jit_create_ctx(sre_pool_t *pool, sre_program_t *prog)
{
    return sre_vm_thompson_create_ctx(pool, prog);
}

SRE_API sre_vm_thompson_exec_pt
sre_vm_thompson_jit_get_handler(sre_vm_thompson_code_t *code)
{
    (void) code;
    return sre_vm_thompson_exec;
}

SRE_API sre_int_t
sre_vm_thompson_jit_free(sre_vm_thompson_code_t *code)
{
    (void) code;
    return SRE_OK;
}

/* ========================================================================
 * bench helper: same contract as ref_bench_lines() in ref_shim.c, over this
 * port (used when oracle/_ref is unavailable; kind "port").
 * ====================================================================== */

typedef struct {
    const char    **regexes;
    const int      *flags;
    int             nregexes, engine, failed;
    const uint8_t  *buf;
    size_t          pitch, linelen, first, count, ovec_slots;
    int32_t        *rc;
    int64_t        *ovec;
    int             reps;
    pthread_barrier_t *ready, *done;
} oj_job_t;

static void *
oj_worker(void *arg)
{
    oj_job_t       *j = arg;
    sre_pool_t     *ppool = sre_create_pool(4096), *pool;
    sre_regex_t    *re;
    sre_program_t  *prog;
    sre_uint_t      ncaps;
    sre_int_t       err_offset, err_id, rc, *ov;
    size_t          i, k, nslots;
    int             rep;

    if (j->nregexes == 1) {
        re = sre_regex_parse(ppool, (sre_char *) j->regexes[0], &ncaps,
                             j->flags ? j->flags[0] : 0, &err_offset);
    } else {
        re = sre_regex_parse_multi(ppool, (sre_char **) j->regexes,
                                   j->nregexes, &ncaps, (int *) j->flags,
                                   &err_offset, &err_id);
    }
    prog = re ? sre_regex_compile(ppool, re) : NULL;
    if (prog == NULL) {
        j->failed = 1;
    }
    /* set-up is outside the timed region: the clock runs between the barriers */
    pthread_barrier_wait(j->ready);
    if (j->failed) {
        pthread_barrier_wait(j->done);
        return NULL;
    }
    nslots = 2 * (ncaps + 1);
    ov = malloc(nslots * sizeof(sre_int_t));
    for (rep = 0; rep < j->reps; rep++)
    for (i = j->first; i < j->first + j->count; i++) {
        const uint8_t *line = j->buf + i * j->pitch;
        pool = sre_create_pool(1024);
        if (j->engine == 2) {
            sre_vm_pike_ctx_t *ctx = sre_vm_pike_create_ctx(pool, prog, ov,
                                         nslots * sizeof(sre_int_t));
            rc = sre_vm_pike_exec(ctx, (sre_char *) line, j->linelen, 1, NULL);
            if (j->ovec) {
                for (k = 0; k < j->ovec_slots; k++) {
                    j->ovec[i * j->ovec_slots + k] =
                        (rc >= 0 && k < nslots) ? ov[k] : -1;
                }
            }
        } else {
            sre_vm_thompson_ctx_t *ctx = sre_vm_thompson_create_ctx(pool, prog);
            rc = sre_vm_thompson_exec(ctx, (sre_char *) line, j->linelen, 1);
        }
        j->rc[i] = (int32_t) rc;
        sre_destroy_pool(pool);
    }
    pthread_barrier_wait(j->done);
    free(ov);
    sre_destroy_pool(ppool);
    return NULL;
}

SRE_API double
ref_bench_lines_reps(const char **regexes, const int *flags, int nregexes,
    int engine, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int nthreads, int reps, int32_t *rc, int64_t *ovec,
    size_t ovec_slots)
{
    pthread_t        *th = calloc(nthreads, sizeof(pthread_t));
    oj_job_t         *jobs = calloc(nthreads, sizeof(oj_job_t));
    pthread_barrier_t ready, done;
    struct timespec   t0, t1;
    size_t            per = (nlines + nthreads - 1) / nthreads;
    int               t, failed = 0;

    pthread_barrier_init(&ready, NULL, nthreads + 1);
    pthread_barrier_init(&done, NULL, nthreads + 1);
    for (t = 0; t < nthreads; t++) {
        oj_job_t *j = &jobs[t];
        j->regexes = regexes; j->flags = flags; j->nregexes = nregexes;
        j->engine = engine; j->buf = buf; j->pitch = pitch;
        j->linelen = linelen;
        j->first = (size_t) t * per;
        j->count = j->first >= nlines ? 0
                   : (j->first + per > nlines ? nlines - j->first : per);
        j->rc = rc; j->ovec = ovec; j->ovec_slots = ovec_slots;
        j->reps = reps; j->ready = &ready; j->done = &done;
        pthread_create(&th[t], NULL, oj_worker, j);
    }
    pthread_barrier_wait(&ready);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&done);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        failed |= jobs[t].failed;
    }
    pthread_barrier_destroy(&ready);
    pthread_barrier_destroy(&done);
    free(th); free(jobs);
    return failed ? -1.0
                  : (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

SRE_API double
ref_bench_lines(const char **regexes, const int *flags, int nregexes,
    int engine, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int nthreads, int32_t *rc, int64_t *ovec,
    size_t ovec_slots)
{
    return ref_bench_lines_reps(regexes, flags, nregexes, engine, buf, nlines,
                                pitch, linelen, nthreads, 1, rc, ovec,
                                ovec_slots);
}

/* one stream in `chunk`-byte calls on one context (see ref_shim.c) */
SRE_API double
ref_bench_stream(const char **regexes, const int *flags, int nregexes,
    int engine, const uint8_t *buf, size_t len, size_t chunk, int reps,
    int *last_rc, long *last_call)
{
    sre_pool_t      *ppool = sre_create_pool(4096), *pool;
    sre_regex_t     *re;
    sre_program_t   *prog;
    sre_uint_t       ncaps;
    sre_int_t        err_offset, err_id;
    struct timespec  t0, t1;
    int              rep;

    if (nregexes == 1) {
        re = sre_regex_parse(ppool, (sre_char *) regexes[0], &ncaps,
                             flags ? flags[0] : 0, &err_offset);
    } else {
        re = sre_regex_parse_multi(ppool, (sre_char **) regexes, nregexes,
                                   &ncaps, (int *) flags, &err_offset, &err_id);
    }
    prog = re ? sre_regex_compile(ppool, re) : NULL;
    if (prog == NULL || engine == 2) {
        return -1.0;
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (rep = 0; rep < reps; rep++) {
        sre_vm_thompson_ctx_t *ctx;
        size_t                 at = 0;
        long                   call = 0;
        sre_int_t              rc = SRE_AGAIN;

        pool = sre_create_pool(4096);
        ctx = sre_vm_thompson_create_ctx(pool, prog);
        do {
            size_t n = len - at < chunk ? len - at : chunk;
            rc = sre_vm_thompson_exec(ctx, (sre_char *) buf + at, n, at + n >= len);
            at += n;
            call++;
        } while (rc == SRE_AGAIN && at < len);
        *last_rc = (int) rc;
        *last_call = call - 1;
        sre_destroy_pool(pool);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    sre_destroy_pool(ppool);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
