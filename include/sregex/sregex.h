/*
 * sregex.h -- drop-in public API of libsregex_cuda (B200-native sregex).
 *
 * This header declares, symbol for symbol, the C API that the reference
 * exports from src/sregex/sregex.h:46-171 (types :46-61, status codes :65-72,
 * pool :82-84, parser :91-107, compiler :118-120, Pike VM :130-134, Thompson VM
 * :144-148, Thompson "JIT" :155-171), so that a program written against the
 * reference re-links against libsregex_cuda unchanged.  Parsing/compilation run
 * on the host; the *_exec entry points run the match on the GPU (there is no
 * CPU fallback: without a usable CUDA device they return SRE_ERROR).
 *
 * Batch / device-pointer extensions live in <sregex_cuda.h>.
 */
#ifndef SREGEX_B200_SREGEX_H
#define SREGEX_B200_SREGEX_H

#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__) && __GNUC__ >= 4
#  define SRE_API    __attribute__ ((visibility ("default")))
#  define SRE_NOAPI  __attribute__ ((visibility ("hidden")))
#else
#  define SRE_API
#  define SRE_NOAPI
#endif

typedef uint8_t    sre_char;    /* reference sregex.h:46-49  */
typedef intptr_t   sre_int_t;   /* reference sregex.h:52-55  */
typedef uintptr_t  sre_uint_t;  /* reference sregex.h:58-61  */

/* status codes, reference sregex.h:65-72 */
enum {
    SRE_OK       = 0,
    SRE_ERROR    = -1,
    SRE_AGAIN    = -2,
    SRE_BUSY     = -3,
    SRE_DONE     = -4,
    SRE_DECLINED = -5
};

/* memory pool, reference sregex.h:78-84 */
typedef struct sre_pool_s  sre_pool_t;

SRE_API sre_pool_t *sre_create_pool(size_t size);
SRE_API void sre_reset_pool(sre_pool_t *pool);
SRE_API void sre_destroy_pool(sre_pool_t *pool);

/* regex flags, reference sregex.h:91-94 */
enum {
    SRE_REGEX_CASELESS = 1,
    SRE_REGEX_NEWLINE  = 2
};

/* parser, reference sregex.h:97-107 */
typedef struct sre_regex_s  sre_regex_t;

SRE_API sre_regex_t *sre_regex_parse(sre_pool_t *pool, sre_char *src,
    sre_uint_t *ncaps, int flags, sre_int_t *err_offset);
SRE_API void sre_regex_dump(sre_regex_t *re);
SRE_API sre_regex_t *sre_regex_parse_multi(sre_pool_t *pool,
    sre_char **regexes, sre_int_t nregexes, sre_uint_t *max_ncaps,
    int *multi_flags, sre_int_t *err_offset, sre_int_t *err_regex_id);

/* compiler, reference sregex.h:113-120 */
typedef struct sre_program_s  sre_program_t;

SRE_API void sre_program_dump(sre_program_t *prog);
SRE_API sre_program_t *sre_regex_compile(sre_pool_t *pool, sre_regex_t *re);

/* Pike VM, reference sregex.h:126-134 */
typedef struct sre_vm_pike_ctx_s  sre_vm_pike_ctx_t;

SRE_API sre_vm_pike_ctx_t *sre_vm_pike_create_ctx(sre_pool_t *pool,
    sre_program_t *prog, sre_int_t *ovector, size_t ovecsize);
SRE_API sre_int_t sre_vm_pike_exec(sre_vm_pike_ctx_t *ctx, sre_char *input,
    size_t len, unsigned eof, sre_int_t **pending_matched);

/* Thompson VM, reference sregex.h:140-148 */
typedef struct sre_vm_thompson_ctx_s  sre_vm_thompson_ctx_t;

SRE_API sre_vm_thompson_ctx_t *sre_vm_thompson_create_ctx(sre_pool_t *pool,
    sre_program_t *prog);
SRE_API sre_int_t sre_vm_thompson_exec(sre_vm_thompson_ctx_t *ctx,
    sre_char *input, size_t len, unsigned eof);

/*
 * Thompson "JIT", reference sregex.h:154-171.  The reference emits x86-64
 * code; here "compiling" means lowering the program to GPU tables and
 * determinising it when it stays small.  The handler has the reference ABI.
 */
typedef struct sre_vm_thompson_code_s  sre_vm_thompson_code_t;

typedef sre_int_t (*sre_vm_thompson_exec_pt)(sre_vm_thompson_ctx_t *ctx,
    sre_char *input, size_t size, unsigned eof);

SRE_API sre_int_t sre_vm_thompson_jit_compile(sre_pool_t *pool,
    sre_program_t *prog, sre_vm_thompson_code_t **pcode);
SRE_API sre_vm_thompson_ctx_t *sre_vm_thompson_jit_create_ctx(sre_pool_t *pool,
    sre_program_t *prog);
SRE_API sre_vm_thompson_exec_pt
    sre_vm_thompson_jit_get_handler(sre_vm_thompson_code_t *code);
SRE_API sre_int_t sre_vm_thompson_jit_free(sre_vm_thompson_code_t *code);

#ifdef __cplusplus
}
#endif

#endif /* SREGEX_B200_SREGEX_H */
