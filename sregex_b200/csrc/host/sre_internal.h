/*
 * sre_internal.h -- host-side data model shared by the front end (parser,
 * compiler), the lowering pass and the test oracle.
 *
 * The *behaviour* follows the reference (AST node kinds: sre_regex.h:17-31,
 * assertion bits: sre_regex.h:34-53, opcodes: sre_vm_bytecode.h:17-27,
 * program header: sre_vm_bytecode.h:72-87) but the layout is our own: the
 * program is a flat, pointer-free, read-only array of 16-byte instructions
 * with branch targets stored as indices, so it can be uploaded to the GPU
 * verbatim and shared by any number of contexts (the reference stores mutable
 * dedup tags inside the program, sre_vm_thompson.c:284 / sre_vm_pike.c:792,
 * which makes it non re-entrant; we keep all run-time state in the contexts).
 */
#ifndef SRE_INTERNAL_H
#define SRE_INTERNAL_H

#include <sregex/sregex.h>
#include <string.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- pool ---------------------------------------------------------------- */

typedef void (*sre_pool_cleanup_pt)(void *data);

SRE_NOAPI void *sre_palloc(sre_pool_t *pool, size_t size);   /* 16B aligned   */
SRE_NOAPI void *sre_pcalloc(sre_pool_t *pool, size_t size);  /* zero filled   */
/* run handler(data) when the pool is reset or destroyed (device buffers) */
SRE_NOAPI int sre_pool_add_cleanup(sre_pool_t *pool,
    sre_pool_cleanup_pt handler, void *data);

/* ---- AST ----------------------------------------------------------------- */

typedef enum {
    SRE_REGEX_TYPE_NIL = 0, SRE_REGEX_TYPE_ALT, SRE_REGEX_TYPE_CAT,
    SRE_REGEX_TYPE_LIT, SRE_REGEX_TYPE_DOT, SRE_REGEX_TYPE_PAREN,
    SRE_REGEX_TYPE_QUEST, SRE_REGEX_TYPE_STAR, SRE_REGEX_TYPE_PLUS,
    SRE_REGEX_TYPE_CLASS, SRE_REGEX_TYPE_NCLASS, SRE_REGEX_TYPE_ASSERT,
    SRE_REGEX_TYPE_TOPLEVEL
} sre_regex_type_t;

/* assertion bits: same numeric values as the reference (sre_regex.h:34-41) */
enum {
    SRE_REGEX_ASSERT_SMALL_Z = 0x01,    /* \z */
    SRE_REGEX_ASSERT_DOLLAR  = 0x02,    /* $  */
    SRE_REGEX_ASSERT_BIG_B   = 0x04,    /* \B */
    SRE_REGEX_ASSERT_SMALL_B = 0x08,    /* \b */
    SRE_REGEX_ASSERT_BIG_A   = 0x10,    /* \A */
    SRE_REGEX_ASSERT_CARET   = 0x20     /* ^  */
};

typedef struct sre_regex_range_s  sre_regex_range_t;
struct sre_regex_range_s {
    sre_char             from, to;
    sre_regex_range_t   *next;
};

struct sre_regex_s {
    sre_regex_type_t     type;
    sre_regex_t         *left, *right;
    sre_uint_t           nregexes;
    union {
        sre_char             ch;
        sre_regex_range_t   *range;
        sre_uint_t          *multi_ncaps;
        sre_uint_t           group;
        sre_uint_t           assertion;
        sre_uint_t           greedy;
        sre_int_t            regex_id;
    } data;
};

/* ---- program ------------------------------------------------------------- */

enum {
    SRE_OPCODE_CHAR = 1, SRE_OPCODE_MATCH = 2, SRE_OPCODE_JMP = 3,
    SRE_OPCODE_SPLIT = 4, SRE_OPCODE_ANY = 5, SRE_OPCODE_SAVE = 6,
    SRE_OPCODE_IN = 7, SRE_OPCODE_NOTIN = 8, SRE_OPCODE_ASSERT = 9
};

typedef struct { sre_char from, to; } sre_vm_range_t;

/* 16 bytes, no pointers: uploaded to the GPU as is */
typedef struct {
    uint8_t     opcode;
    uint8_t     ch;         /* CHAR                                          */
    uint16_t    nranges;    /* IN / NOTIN                                    */
    int32_t     x;          /* JMP target / SPLIT preferred branch           */
    int32_t     y;          /* SPLIT second branch                           */
    int32_t     v;          /* SAVE slot | ASSERT kind | MATCH regex id |
                               IN/NOTIN first range index                    */
} sre_instruction_t;

#define SRE_PROGRAM_MAGIC  0x53524550u  /* "SREP" */

struct sre_program_s {
    uint32_t             magic;
    uint32_t             len;
    sre_instruction_t   *insts;
    sre_vm_range_t      *ranges;
    uint32_t             nranges;
    uint32_t             nullable;
    int32_t              leading_byte;      /* -1 if not a single CHAR      */
    uint32_t             nleading;          /* leading consuming instrs     */
    int32_t             *leading;           /* their pcs (Pike prefilter)   */
    sre_uint_t           ovecsize;          /* bytes, all regexes' groups   */
    sre_uint_t           nregexes;
    sre_uint_t          *multi_ncaps;
    void                *lowered;           /* cache owned by the CUDA side */
    sre_pool_t          *pool;
};

SRE_NOAPI sre_regex_t *sre_regex_create(sre_pool_t *pool, sre_regex_type_t type,
    sre_regex_t *left, sre_regex_t *right);
/* text of sre_program_dump(), malloc'ed; used by the parity tests */
SRE_API char *sre_program_dump_str(sre_program_t *prog);

#define sre_isword(c)                                                        \
    (((c) >= '0' && (c) <= '9') || ((c) >= 'A' && (c) <= 'Z')                \
     || ((c) >= 'a' && (c) <= 'z') || (c) == '_')   /* sre_core.h:31-35 */

#ifdef __cplusplus
}
#endif
#endif
