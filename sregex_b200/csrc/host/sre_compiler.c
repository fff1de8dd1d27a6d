/*
 * sre_compiler.c -- AST -> bytecode: sre_regex_compile / sre_program_dump.
 *
 * Produces the same instruction sequence as the reference's
 * sre_regex_emit_bytecode (sre_regex_compiler.c:288-482: ALT :296-317, PAREN
 * :367-382, QUEST :384-402, STAR :404-426, PLUS :428-448, TOPLEVEL :457-470),
 * the same leading-byte analysis for the Pike prefilter (:123-241) and the same
 * dump text (sre_vm_bytecode.c:14-128).  Unlike the reference the program is a
 * flat array with index branch targets and no mutable fields, see
 * sre_internal.h.
 */
#define _GNU_SOURCE
#include "sre_internal.h"

/* ---- sizing -------------------------------------------------------------- */

typedef struct { uint32_t ninsts, nranges; } sre_size_t;

static void
measure(sre_regex_t *r, sre_size_t *sz)
{
    sre_regex_range_t *range;

    switch (r->type) {
    case SRE_REGEX_TYPE_ALT:
        sz->ninsts += 2;    /* split + jmp */
        measure(r->left, sz);
        measure(r->right, sz);
        break;
    case SRE_REGEX_TYPE_CAT:
        measure(r->left, sz);
        measure(r->right, sz);
        break;
    case SRE_REGEX_TYPE_CLASS:
    case SRE_REGEX_TYPE_NCLASS:
        for (range = r->data.range; range; range = range->next) {
            sz->nranges++;
        }
        /* fall through */
    case SRE_REGEX_TYPE_LIT:
    case SRE_REGEX_TYPE_DOT:
    case SRE_REGEX_TYPE_ASSERT:
        sz->ninsts += 1;
        break;
    case SRE_REGEX_TYPE_PAREN:      /* save, save */
    case SRE_REGEX_TYPE_STAR:       /* split, jmp */
        sz->ninsts += 2;
        measure(r->left, sz);
        break;
    case SRE_REGEX_TYPE_QUEST:      /* split */
    case SRE_REGEX_TYPE_PLUS:       /* split */
    case SRE_REGEX_TYPE_TOPLEVEL:   /* match */
        sz->ninsts += 1;
        measure(r->left, sz);
        break;
    default:                        /* NIL */
        break;
    }
}

/* ---- emission ------------------------------------------------------------ */

typedef struct {
    sre_instruction_t  *insts;
    sre_vm_range_t     *ranges;
    uint32_t            pc, nranges;
} sre_emitter_t;

static sre_instruction_t *
emit_op(sre_emitter_t *e, int opcode)
{
    sre_instruction_t *in = &e->insts[e->pc++];
    in->opcode = (uint8_t) opcode;
    return in;
}

/* x = preferred branch */
static void
set_split(sre_instruction_t *in, int32_t body, int32_t skip, unsigned greedy)
{
    in->x = greedy ? body : skip;
    in->y = greedy ? skip : body;
}

static void
emit(sre_emitter_t *e, sre_regex_t *r)
{
    sre_instruction_t  *split, *jmp, *in;
    sre_regex_range_t  *range;
    int32_t             body;

    switch (r->type) {
    case SRE_REGEX_TYPE_ALT:
        split = emit_op(e, SRE_OPCODE_SPLIT);
        split->x = e->pc;
        emit(e, r->left);
        jmp = emit_op(e, SRE_OPCODE_JMP);
        split->y = e->pc;
        emit(e, r->right);
        jmp->x = e->pc;
        break;

    case SRE_REGEX_TYPE_CAT:
        emit(e, r->left);
        emit(e, r->right);
        break;

    case SRE_REGEX_TYPE_LIT:
        emit_op(e, SRE_OPCODE_CHAR)->ch = r->data.ch;
        break;

    case SRE_REGEX_TYPE_DOT:
        emit_op(e, SRE_OPCODE_ANY);
        break;

    case SRE_REGEX_TYPE_CLASS:
    case SRE_REGEX_TYPE_NCLASS:
        in = emit_op(e, r->type == SRE_REGEX_TYPE_CLASS ? SRE_OPCODE_IN
                                                        : SRE_OPCODE_NOTIN);
        in->v = e->nranges;
        for (range = r->data.range; range; range = range->next) {
            e->ranges[e->nranges].from = range->from;
            e->ranges[e->nranges].to = range->to;
            e->nranges++;
            in->nranges++;
        }
        break;

    case SRE_REGEX_TYPE_PAREN:
        emit_op(e, SRE_OPCODE_SAVE)->v = 2 * r->data.group;
        emit(e, r->left);
        emit_op(e, SRE_OPCODE_SAVE)->v = 2 * r->data.group + 1;
        break;

    case SRE_REGEX_TYPE_QUEST:
        split = emit_op(e, SRE_OPCODE_SPLIT);
        body = e->pc;
        emit(e, r->left);
        set_split(split, body, e->pc, r->data.greedy);
        break;

    case SRE_REGEX_TYPE_STAR:
        split = emit_op(e, SRE_OPCODE_SPLIT);
        body = e->pc;
        emit(e, r->left);
        emit_op(e, SRE_OPCODE_JMP)->x = (int32_t) (split - e->insts);
        set_split(split, body, e->pc, r->data.greedy);
        break;

    case SRE_REGEX_TYPE_PLUS:
        body = e->pc;
        emit(e, r->left);
        split = emit_op(e, SRE_OPCODE_SPLIT);
        set_split(split, body, e->pc, r->data.greedy);
        break;

    case SRE_REGEX_TYPE_ASSERT:
        emit_op(e, SRE_OPCODE_ASSERT)->v = (int32_t) r->data.assertion;
        break;

    case SRE_REGEX_TYPE_TOPLEVEL:
        emit(e, r->left);
        emit_op(e, SRE_OPCODE_MATCH)->v = (int32_t) r->data.regex_id;
        break;

    default:    /* NIL emits nothing */
        break;
    }
}

/* ---- leading bytes (Pike prefilter) --------------------------------------
 * DFS from pc 0 through SPLIT(x,y)/JMP/SAVE/ASSERT, skipping the `any` of the
 * ".*?" prefix (pc 1).  Collects the first consuming instructions; a reachable
 * MATCH makes the program nullable and an ANY makes the set useless
 * (behaviour of sre_regex_compiler.c:123-241, incl. the early stop after
 * either of those).  Returns SRE_OK / SRE_DONE (nullable) / SRE_DECLINED.   */

static sre_int_t
leading_walk(sre_program_t *prog, int32_t pc, uint8_t *seen)
{
    sre_instruction_t  *in, *other;
    sre_int_t           rc;
    uint32_t            i;

    for (;;) {
        if ((uint32_t) pc >= prog->len || seen[pc] || pc == 1) {
            return SRE_OK;
        }
        seen[pc] = 1;
        in = &prog->insts[pc];

        switch (in->opcode) {
        case SRE_OPCODE_SPLIT:
            rc = leading_walk(prog, in->x, seen);
            if (rc != SRE_OK) {
                return rc;
            }
            pc = in->y;
            continue;
        case SRE_OPCODE_JMP:
            pc = in->x;
            continue;
        case SRE_OPCODE_SAVE:
        case SRE_OPCODE_ASSERT:
            pc++;
            continue;
        case SRE_OPCODE_MATCH:
            prog->nullable = 1;
            return SRE_DONE;
        case SRE_OPCODE_ANY:
            return SRE_DECLINED;
        default:
            /* CHAR / IN / NOTIN; identical CHARs are listed once */
            if (in->opcode == SRE_OPCODE_CHAR) {
                for (i = 0; i < prog->nleading; i++) {
                    other = &prog->insts[prog->leading[i]];
                    if (other->opcode == SRE_OPCODE_CHAR
                        && other->ch == in->ch)
                    {
                        return SRE_OK;
                    }
                }
            }
            prog->leading[prog->nleading++] = pc;
            return SRE_OK;
        }
    }
}

SRE_API sre_program_t *
sre_regex_compile(sre_pool_t *pool, sre_regex_t *re)
{
    sre_program_t  *prog;
    sre_emitter_t   e;
    sre_size_t      sz = { 0, 0 };
    sre_uint_t      i;
    uint8_t        *seen;
    sre_int_t       rc;

    if (re == NULL || re->nregexes == 0) {
        return NULL;
    }
    measure(re, &sz);

    prog = sre_pcalloc(pool, sizeof(sre_program_t));
    if (prog == NULL) {
        return NULL;
    }
    prog->insts = sre_pcalloc(pool, (sz.ninsts + 1) * sizeof(sre_instruction_t));
    prog->ranges = sre_pcalloc(pool, (sz.nranges + 1) * sizeof(sre_vm_range_t));
    prog->leading = sre_pcalloc(pool, (sz.ninsts + 1) * sizeof(int32_t));
    prog->multi_ncaps = sre_palloc(pool, re->nregexes * sizeof(sre_uint_t));
    seen = sre_pcalloc(pool, sz.ninsts + 1);
    if (!prog->insts || !prog->ranges || !prog->leading || !prog->multi_ncaps
        || !seen)
    {
        return NULL;
    }

    e.insts = prog->insts;
    e.ranges = prog->ranges;
    e.pc = 0;
    e.nranges = 0;
    emit(&e, re);
    if (e.pc != sz.ninsts || e.nranges != sz.nranges) {
        return NULL;
    }

    prog->magic = SRE_PROGRAM_MAGIC;
    prog->len = e.pc;
    prog->nranges = e.nranges;
    prog->pool = pool;
    prog->nregexes = re->nregexes;
    memcpy(prog->multi_ncaps, re->data.multi_ncaps,
           re->nregexes * sizeof(sre_uint_t));

    /* one capture vector holds every regex's groups back to back */
    prog->ovecsize = 0;
    for (i = 0; i < prog->nregexes; i++) {
        prog->ovecsize += prog->multi_ncaps[i] + 1;
    }
    prog->ovecsize *= 2 * sizeof(sre_int_t);

    prog->leading_byte = -1;
    rc = leading_walk(prog, 0, seen);
    if (rc != SRE_OK || prog->nullable) {
        prog->nleading = 0;
    }
    if (prog->nleading == 1
        && prog->insts[prog->leading[0]].opcode == SRE_OPCODE_CHAR)
    {
        prog->leading_byte = prog->insts[prog->leading[0]].ch;
    }
    return prog;
}

/* ---- dump ---------------------------------------------------------------- */

static void
dump_instruction(FILE *f, sre_program_t *prog, uint32_t pc)
{
    sre_instruction_t  *in = &prog->insts[pc];
    const char         *s;
    uint32_t            i;

    fprintf(f, "%2d. ", (int) pc);
    switch (in->opcode) {
    case SRE_OPCODE_SPLIT:
        fprintf(f, "split %d, %d", in->x, in->y);
        break;
    case SRE_OPCODE_JMP:
        fprintf(f, "jmp %d", in->x);
        break;
    case SRE_OPCODE_CHAR:
        fprintf(f, "char %d", (int) in->ch);
        break;
    case SRE_OPCODE_IN:
    case SRE_OPCODE_NOTIN:
        fprintf(f, in->opcode == SRE_OPCODE_IN ? "in" : "notin");
        for (i = 0; i < in->nranges; i++) {
            fprintf(f, "%s %d-%d", i ? "," : "", prog->ranges[in->v + i].from,
                    prog->ranges[in->v + i].to);
        }
        break;
    case SRE_OPCODE_ANY:
        fprintf(f, "any");
        break;
    case SRE_OPCODE_MATCH:
        fprintf(f, "match %d", in->v);
        break;
    case SRE_OPCODE_SAVE:
        fprintf(f, "save %d", in->v);
        break;
    case SRE_OPCODE_ASSERT:
        switch (in->v) {
        case SRE_REGEX_ASSERT_BIG_A:   s = "\\A"; break;
        case SRE_REGEX_ASSERT_CARET:   s = "^";   break;
        case SRE_REGEX_ASSERT_SMALL_Z: s = "\\z"; break;
        case SRE_REGEX_ASSERT_BIG_B:   s = "\\B"; break;
        case SRE_REGEX_ASSERT_SMALL_B: s = "\\b"; break;
        case SRE_REGEX_ASSERT_DOLLAR:  s = "$";   break;
        default:                       s = "?";   break;
        }
        fprintf(f, "assert %s", s);
        break;
    default:
        fprintf(f, "unknown");
        break;
    }
    fputc('\n', f);
}

SRE_API void
sre_program_dump(sre_program_t *prog)
{
    uint32_t pc;
    for (pc = 0; pc < prog->len; pc++) {
        dump_instruction(stdout, prog, pc);
    }
}

SRE_API char *
sre_program_dump_str(sre_program_t *prog)
{
    char    *buf = NULL;
    size_t   len = 0;
    FILE    *f = open_memstream(&buf, &len);
    uint32_t pc;

    if (f == NULL) {
        return NULL;
    }
    for (pc = 0; pc < prog->len; pc++) {
        dump_instruction(f, prog, pc);
    }
    fclose(f);
    return buf;
}
