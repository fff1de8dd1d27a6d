/*
 * sregex-cli.c -- command-line driver with the interface and output format of
 * the reference's test CLI (src/sre_cli.c), so that the reference's own harness
 * (t/SRegex.pm:73-84 pipes "<len>\n<subject>" to `sregex-cli --stdin [--flags f]
 * [-n k] re...` and parses the six result lines, :293-441) can drive ANY library
 * exporting the sregex API.  Written from that interface description; it only
 * uses the public API of <sregex/sregex.h>, which is the point: the same source
 * links against libsregex_cuda (GPU), the CPU oracle, or the reference.
 *
 *   usage: sregex-cli [--stdin] [--flags "i i ..."] [-n nregexes] re... [subject...]
 *
 * Per subject it prints:
 *   ## <subject> (len N)
 *   thompson <verdict>
 *   splitted thompson <verdict>
 *   jitted thompson <verdict> | jitted thompson disabled
 *   splitted jitted thompson <verdict> | splitted jitted thompson disabled
 *   pike match <id> (s, e) ... | pike no match
 *   splitted pike [(s, e)](ps, pe) ... match <id> (s, e) ... | ... no match
 * where "splitted" feeds an empty chunk before every 1-byte chunk and an empty
 * eof chunk at the end.
 */
#include <sregex/sregex.h>
#include <stdio.h>
#include <string.h>

typedef sre_int_t (*thompson_fn)(sre_vm_thompson_ctx_t *, sre_char *, size_t, unsigned);

static const char *
verdict(sre_int_t rc)
{
    switch (rc) {
    case SRE_OK:       return "match";
    case SRE_DECLINED: return "no match";
    case SRE_AGAIN:    return "again";
    case SRE_ERROR:    return "error";
    default:           return "unknown";
    }
}

/* the reference CLI's chunking: (empty, byte)* then an empty eof chunk */
static sre_int_t
feed_thompson(thompson_fn exec, sre_vm_thompson_ctx_t *ctx, sre_char *s, size_t len, int split)
{
    sre_int_t  rc;
    sre_char   one;
    size_t     i;

    if (!split) {
        return exec(ctx, s, len, 1);
    }
    for (i = 0; i < len; i++) {
        rc = exec(ctx, NULL, 0, 0);
        if (rc != SRE_AGAIN) {
            return rc;
        }
        one = s[i];
        rc = exec(ctx, &one, 1, 0);
        if (rc != SRE_AGAIN) {
            return rc;
        }
    }
    return exec(ctx, NULL, 0, 1);
}

static void
print_pike_result(sre_int_t rc, sre_int_t *ovector, sre_uint_t ncaps)
{
    sre_uint_t i;

    if (rc >= 0) {
        printf("match %ld", (long) rc);
        for (i = 0; i < 2 * (ncaps + 1); i += 2) {
            printf(" (%ld, %ld)", (long) ovector[i], (long) ovector[i + 1]);
        }
        printf("\n");
    } else if (rc == SRE_AGAIN) {
        printf("again\n");
    } else if (rc == SRE_DECLINED) {
        printf("no match\n");
    } else if (rc == SRE_ERROR) {
        printf("error\n");
    } else {
        printf("unknown (%d)\n", (int) rc);
    }
}

static void
run_subject(sre_program_t *prog, sre_char *s, size_t len, sre_int_t *ovector, size_t ovecsize,
    sre_uint_t ncaps)
{
    sre_pool_t               *pool = sre_create_pool(1024);
    sre_vm_thompson_ctx_t    *tctx;
    sre_vm_thompson_code_t   *code = NULL;
    sre_vm_pike_ctx_t        *pctx;
    sre_int_t                *pending, rc;
    sre_char                  one;
    size_t                    i;
    int                       split;

    printf("## %.*s (len %d)\n", (int) len, s, (int) len);

    for (split = 0; split < 2; split++) {
        printf("%sthompson ", split ? "splitted " : "");
        tctx = sre_vm_thompson_create_ctx(pool, prog);
        printf("%s\n", tctx ? verdict(feed_thompson(sre_vm_thompson_exec, tctx, s, len, split)) : "error");
        sre_reset_pool(pool);
    }

    rc = sre_vm_thompson_jit_compile(pool, prog, &code);
    if (rc == SRE_DECLINED) {
        printf("jitted thompson disabled\n");
        printf("splitted jitted thompson disabled\n");
    } else if (rc != SRE_OK) {
        fprintf(stderr, "failed to run thompson jit compile: %ld\n", (long) rc);
        exit(2);
    } else {
        thompson_fn handler = sre_vm_thompson_jit_get_handler(code);
        for (split = 0; split < 2; split++) {
            printf("%sjitted thompson ", split ? "splitted " : "");
            tctx = sre_vm_thompson_jit_create_ctx(pool, prog);
            printf("%s\n", tctx ? verdict(feed_thompson(handler, tctx, s, len, split)) : "error");
        }
        sre_vm_thompson_jit_free(code);
        sre_reset_pool(pool);
    }

    printf("pike ");
    pctx = sre_vm_pike_create_ctx(pool, prog, ovector, ovecsize);
    print_pike_result(pctx ? sre_vm_pike_exec(pctx, s, len, 1, NULL) : SRE_ERROR, ovector, ncaps);
    sre_reset_pool(pool);

    printf("splitted pike ");
    pctx = sre_vm_pike_create_ctx(pool, prog, ovector, ovecsize);
    rc = pctx ? SRE_AGAIN : SRE_ERROR;
    for (i = 0; i < len && rc == SRE_AGAIN; i++) {
        rc = sre_vm_pike_exec(pctx, NULL, 0, 0, NULL);
        if (rc != SRE_AGAIN) {
            break;
        }
        one = s[i];
        pending = NULL;
        rc = sre_vm_pike_exec(pctx, &one, 1, 0, &pending);
        if (rc == SRE_AGAIN) {
            /* temporary captures, then the pending match if there is one */
            printf("[(%ld, %ld)]", (long) ovector[0], (long) ovector[1]);
            if (pending) {
                printf("(%ld, %ld) ", (long) pending[0], (long) pending[1]);
            } else {
                printf(" ");
            }
        }
    }
    if (rc == SRE_AGAIN) {
        pending = NULL;
        rc = sre_vm_pike_exec(pctx, NULL, 0, 1, &pending);
    }
    print_pike_result(rc, ovector, ncaps);

    sre_destroy_pool(pool);
}

static int
parse_flags(const char *str, int nregexes, int *flags)
{
    int i = 0;

    for (; *str; str++) {
        if (i >= nregexes) {
            fprintf(stderr, "Too many flags given but only %d regexes specified.\n", nregexes);
            return -1;
        }
        if (*str == ' ') {
            i++;
        } else if (*str == 'i') {
            flags[i] |= SRE_REGEX_CASELESS;
        } else {
            fprintf(stderr, "Bad regex flag '%c' for regex %d\n", *str, i);
            return -1;
        }
    }
    return 0;
}

int
main(int argc, char **argv)
{
    const char      *flags_str = NULL;
    int              from_stdin = 0, nregexes = 1, i, *flags = NULL;
    sre_pool_t      *ppool, *cpool;
    sre_regex_t     *re;
    sre_program_t   *prog;
    sre_uint_t       ncaps;
    sre_int_t        err_offset, err_id, *ovector;
    size_t           ovecsize;

    if (argc < 2) {
        fprintf(stderr, "usage: sregex-cli regexp string...\n       sregex-cli --stdin regexp\n");
        return 2;
    }
    for (i = 1; i < argc && argv[i][0] == '-' && argv[i][1] != '\0'; i++) {
        if (strcmp(argv[i], "--stdin") == 0) {
            from_stdin = 1;
        } else if (strcmp(argv[i], "--flags") == 0 && i + 1 < argc) {
            flags_str = argv[++i];
        } else if (strcmp(argv[i], "-n") == 0 && i + 1 < argc) {
            nregexes = atoi(argv[++i]);
            if (nregexes <= 0) {
                fprintf(stderr, "invalid -n value: %s.\n", argv[i]);
                return 1;
            }
        } else {
            fprintf(stderr, "unknown option: %s\n", argv[i]);
            return 1;
        }
    }
    if (argc - i < nregexes) {
        fprintf(stderr, "at least %d regexes should be specified\n", nregexes);
        return 1;
    }
    if (flags_str) {
        flags = calloc(nregexes, sizeof(int));
        if (flags == NULL || parse_flags(flags_str, nregexes, flags) != 0) {
            fprintf(stderr, "Bad --flags option value: %s", flags_str);
            return 1;
        }
    }

    ppool = sre_create_pool(1024);
    if (nregexes == 1) {
        re = sre_regex_parse(ppool, (sre_char *) argv[i], &ncaps, flags ? flags[0] : 0, &err_offset);
        if (re == NULL) {
            if (err_offset >= 0) {
                fprintf(stderr, "[error] syntax error at pos %lld\n", (long long) err_offset);
            } else {
                fprintf(stderr, "unknown error\n");
            }
            return 1;
        }
    } else {
        re = sre_regex_parse_multi(ppool, (sre_char **) &argv[i], nregexes, &ncaps, flags, &err_offset,
                                   &err_id);
        if (re == NULL) {
            if (err_offset >= 0) {
                fprintf(stderr, "[error] regex %lu: syntax error at pos %ld\n", (unsigned long) err_id,
                        (long) err_offset);
            } else {
                fprintf(stderr, "unknown error\n");
            }
            return 1;
        }
    }
    i += nregexes;

    sre_regex_dump(re);
    printf("\n");
    printf("captures: %ld\n", (long) ncaps);

    cpool = sre_create_pool(1024);
    prog = sre_regex_compile(cpool, re);
    if (prog == NULL) {
        fprintf(stderr, "failed to compile the regex.\n");
        return 2;
    }
    sre_destroy_pool(ppool);
    sre_program_dump(prog);

    ovecsize = 2 * (ncaps + 1) * sizeof(sre_int_t);
    ovector = malloc(ovecsize);
    if (ovector == NULL) {
        return 2;
    }

    if (from_stdin) {
        int n;
        while (scanf("%d", &n) == 1) {
            sre_char *s;
            if (getchar() != '\n') {
                fprintf(stderr, "the next character after the chunk size must be a newline");
                return 1;
            }
            s = malloc(n ? n : 1);
            if (s == NULL || fread(s, 1, n, stdin) < (size_t) n) {
                fprintf(stderr, "failed to read %d bytes of string from stdin.", n);
                return 2;
            }
            run_subject(prog, s, n, ovector, ovecsize, ncaps);
            free(s);
        }
    } else {
        if (i >= argc) {
            fprintf(stderr, "no subject string specified.\n");
            return 1;
        }
        for (; i < argc; i++) {
            run_subject(prog, (sre_char *) argv[i], strlen(argv[i]), ovector, ovecsize, ncaps);
        }
    }

    sre_destroy_pool(cpool);
    free(ovector);
    free(flags);
    return 0;
}
