/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into libsregex_cuda).
 *
 * Thin helpers that are compiled *together with the unmodified reference
 * sources* (read where they lie under /root/reference, see oracle/Makefile)
 * into oracle/_ref/libsregex_ref.so.  The reference's public API
 * (src/sregex/sregex.h:82-171) is called directly from Python via ctypes;
 * this file only adds what cannot be done through that API:
 *
 *   - ref_program_dump_str(): sre_program_dump() prints to stdout
 *     (sre_vm_bytecode.c:14-27); we need the text in memory, so we call the
 *     reference's own sre_dump_instruction() (sre_vm_bytecode.c:30-128) on an
 *     open_memstream.
 *   - ref_bench_lines()/ref_bench_buffer(): the CPU baseline of SURVEY.md
 *     section 8(d): one OS thread per core, one private compiled program per
 *     thread, fresh ctx + fresh pool per line, wall clock over all threads.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <sregex/sregex.h>
#include <sregex/sre_vm_bytecode.h>

char *
ref_program_dump_str(sre_program_t *prog)
{
    char    *buf = NULL;
    size_t   len = 0;
    FILE    *f = open_memstream(&buf, &len);
    sre_instruction_t *pc, *end = prog->start + prog->len;

    for (pc = prog->start; pc < end; pc++) {
        sre_dump_instruction(f, pc, prog->start);
        fputc('\n', f);
    }
    fclose(f);
    return buf;
}

void ref_free(void *p) { free(p); }

long ref_program_len(sre_program_t *prog) { return (long) prog->len; }
long ref_program_ovecsize(sre_program_t *prog) { return (long) prog->ovecsize; }

enum { REF_ENGINE_THOMPSON = 0, REF_ENGINE_JIT = 1, REF_ENGINE_PIKE = 2 };

typedef struct {
    const char    **regexes;
    const int      *flags;
    int             nregexes;
    int             engine;
    const uint8_t  *buf;
    size_t          pitch;      /* bytes between line starts */
    size_t          linelen;
    size_t          first, count;
    int32_t        *rc;         /* per line */
    int64_t        *ovec;       /* per line, ovec_slots each (pike) or NULL */
    size_t          ovec_slots;
    int             failed;
    int             reps;       /* passes over the lines inside the timed region */
    pthread_barrier_t *ready, *done;
} ref_job_t;

static sre_program_t *
ref_build(sre_pool_t *pool, ref_job_t *j, sre_uint_t *ncaps)
{
    sre_int_t     err_offset, err_id;
    sre_regex_t  *re;

    if (j->nregexes == 1) {
        re = sre_regex_parse(pool, (sre_char *) j->regexes[0], ncaps,
                             j->flags ? j->flags[0] : 0, &err_offset);
    } else {
        re = sre_regex_parse_multi(pool, (sre_char **) j->regexes,
                                   j->nregexes, ncaps, (int *) j->flags,
                                   &err_offset, &err_id);
    }
    if (re == NULL) {
        return NULL;
    }
    return sre_regex_compile(pool, re);
}

static void *
ref_worker(void *arg)
{
    ref_job_t                *j = arg;
    sre_pool_t               *ppool, *pool;
    sre_program_t            *prog;
    sre_uint_t                ncaps;
    sre_vm_thompson_code_t   *code = NULL;
    sre_vm_thompson_exec_pt   handler = NULL;
    size_t                    i, k, nslots;
    sre_int_t                *ov;

    int                       rep;

    /* set-up (parse, compile, JIT) is outside the timed region: the clock runs
     * between the two barriers */
    ppool = sre_create_pool(4096);
    prog = ref_build(ppool, j, &ncaps);
    if (prog == NULL) {
        j->failed = 1;
    } else if (j->engine == REF_ENGINE_JIT) {
        if (sre_vm_thompson_jit_compile(ppool, prog, &code) != SRE_OK) {
            j->failed = 1;
        } else {
            handler = sre_vm_thompson_jit_get_handler(code);
        }
    }
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
        sre_int_t      rc;

        /* a fresh pool per line: sre_reset_pool does not rewind
         * pool->current (sre_palloc.c:118-148), see SURVEY.md 8(a) */
        pool = sre_create_pool(1024);
        if (j->engine == REF_ENGINE_THOMPSON) {
            sre_vm_thompson_ctx_t *ctx = sre_vm_thompson_create_ctx(pool, prog);
            rc = sre_vm_thompson_exec(ctx, (sre_char *) line, j->linelen, 1);
        } else if (j->engine == REF_ENGINE_JIT) {
            sre_vm_thompson_ctx_t *ctx =
                sre_vm_thompson_jit_create_ctx(pool, prog);
            rc = handler(ctx, (sre_char *) line, j->linelen, 1);
        } else {
            sre_vm_pike_ctx_t *ctx = sre_vm_pike_create_ctx(pool, prog, ov,
                                         nslots * sizeof(sre_int_t));
            rc = sre_vm_pike_exec(ctx, (sre_char *) line, j->linelen, 1, NULL);
            if (j->ovec) {
                for (k = 0; k < j->ovec_slots; k++) {
                    j->ovec[i * j->ovec_slots + k] =
                        (rc >= 0 && k < nslots) ? ov[k] : -1;
                }
            }
        }
        j->rc[i] = (int32_t) rc;
        sre_destroy_pool(pool);
    }

    pthread_barrier_wait(j->done);
    free(ov);
    if (code) {
        sre_vm_thompson_jit_free(code);
    }
    sre_destroy_pool(ppool);
    return NULL;
}

/*
 * Runs `engine` over nlines lines of linelen bytes (line i starts at
 * buf + i*pitch) on nthreads OS threads, reps times over.  Returns the wall
 * seconds of the matching itself (thread creation, parse, compile and JIT are
 * done before the clock starts), <0 on error.
 */
double
ref_bench_lines_reps(const char **regexes, const int *flags, int nregexes,
    int engine, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int nthreads, int reps, int32_t *rc, int64_t *ovec,
    size_t ovec_slots)
{
    pthread_t        *th = calloc(nthreads, sizeof(pthread_t));
    ref_job_t        *jobs = calloc(nthreads, sizeof(ref_job_t));
    pthread_barrier_t ready, done;
    struct timespec   t0, t1;
    size_t            per = (nlines + nthreads - 1) / nthreads;
    int               t, failed = 0;

    pthread_barrier_init(&ready, NULL, nthreads + 1);
    pthread_barrier_init(&done, NULL, nthreads + 1);
    for (t = 0; t < nthreads; t++) {
        ref_job_t *j = &jobs[t];
        j->regexes = regexes; j->flags = flags; j->nregexes = nregexes;
        j->engine = engine; j->buf = buf; j->pitch = pitch;
        j->linelen = linelen;
        j->first = (size_t) t * per;
        j->count = j->first >= nlines ? 0
                   : (j->first + per > nlines ? nlines - j->first : per);
        j->rc = rc; j->ovec = ovec; j->ovec_slots = ovec_slots;
        j->reps = reps; j->ready = &ready; j->done = &done;
        pthread_create(&th[t], NULL, ref_worker, j);
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
    if (failed) {
        return -1.0;
    }
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

double
ref_bench_lines(const char **regexes, const int *flags, int nregexes,
    int engine, const uint8_t *buf, size_t nlines, size_t pitch,
    size_t linelen, int nthreads, int32_t *rc, int64_t *ovec,
    size_t ovec_slots)
{
    return ref_bench_lines_reps(regexes, flags, nregexes, engine, buf, nlines,
                                pitch, linelen, nthreads, 1, rc, ovec,
                                ovec_slots);
}

/*
 * One stream fed to ONE context in `chunk`-byte calls with SRE_AGAIN carry
 * (how bench/sregex.c and the nginx module drive a long input; a single
 * stream is sequential on the CPU: one core).  Thompson engines only.
 * *last_rc = rc of the last call made, *last_call = its index.  Returns the
 * seconds of the calls (compile excluded), <0 on error.
 */
double
ref_bench_stream(const char **regexes, const int *flags, int nregexes,
    int engine, const uint8_t *buf, size_t len, size_t chunk, int reps,
    int *last_rc, long *last_call)
{
    ref_job_t                 j;
    sre_pool_t               *ppool, *pool;
    sre_program_t            *prog;
    sre_uint_t                ncaps;
    sre_vm_thompson_code_t   *code = NULL;
    sre_vm_thompson_exec_pt   handler = sre_vm_thompson_exec;
    struct timespec           t0, t1;
    int                       rep;

    memset(&j, 0, sizeof(j));
    j.regexes = regexes; j.flags = flags; j.nregexes = nregexes;
    ppool = sre_create_pool(4096);
    prog = ref_build(ppool, &j, &ncaps);
    if (prog == NULL || engine == REF_ENGINE_PIKE) {
        return -1.0;
    }
    if (engine == REF_ENGINE_JIT) {
        if (sre_vm_thompson_jit_compile(ppool, prog, &code) != SRE_OK) {
            return -1.0;
        }
        handler = sre_vm_thompson_jit_get_handler(code);
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (rep = 0; rep < reps; rep++) {
        sre_vm_thompson_ctx_t *ctx;
        size_t                 at = 0;
        long                   call = 0;
        sre_int_t              rc = SRE_AGAIN;

        pool = sre_create_pool(4096);
        ctx = engine == REF_ENGINE_JIT ? sre_vm_thompson_jit_create_ctx(pool, prog)
                                       : sre_vm_thompson_create_ctx(pool, prog);
        do {
            size_t n = len - at < chunk ? len - at : chunk;
            rc = handler(ctx, (sre_char *) buf + at, n, at + n >= len);
            at += n;
            call++;
        } while (rc == SRE_AGAIN && at < len);
        *last_rc = (int) rc;
        *last_call = call - 1;
        sre_destroy_pool(pool);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (code) {
        sre_vm_thompson_jit_free(code);
    }
    sre_destroy_pool(ppool);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
