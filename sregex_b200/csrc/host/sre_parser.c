/*
 * sre_parser.c -- regex front end: sre_regex_parse / sre_regex_parse_multi.
 *
 * Behavioural spec = the reference's bison grammar and hand lexer
 * (src/sregex/sre_yyparser.y: grammar :105-345, lexer :350-1795, entry points
 * :1806-1986, counted-repetition desugaring :2011-2084, caseless class folding
 * sre_regex.c:170-214).  The implementation is new: a table-driven tokenizer
 * and a recursive-descent parser with one token of look-ahead.  An LALR(1)
 * parser reports a syntax error at the first token that cannot extend a viable
 * prefix, and so does this one, hence *err_offset (start of the offending
 * token, .y:1798-1803) is identical.  Parity is checked by diffing
 * sre_program_dump() text against the reference over every regex of the
 * reference's t/ suite (tests/test_frontend.py).
 */
#include "sre_internal.h"
#include <ctype.h>

/* ---- tokens -------------------------------------------------------------- */

enum {
    /* punctuation tokens are their own character code */
    TOK_CHAR = 256, TOK_EOF, TOK_BAD, TOK_CQUANT, TOK_CLASS, TOK_ASSERT
};

typedef struct {
    int              kind;
    const sre_char  *pos;       /* start of the token (error offset)         */
    sre_char         ch;        /* TOK_CHAR                                  */
    int              from, to;  /* TOK_CQUANT; to == -1: unbounded           */
    sre_regex_t     *re;        /* TOK_CLASS / TOK_ASSERT                    */
} sre_token_t;

typedef struct {
    sre_pool_t      *pool;
    const sre_char  *src;       /* read cursor                               */
    int              flags;
    sre_uint_t      *ncaps;
    sre_token_t      tok;       /* look-ahead                                */
    const sre_char  *err_pos;
    int              oom;
} sre_parser_t;

/* escape -> inclusive byte ranges.  Values per .y:361-384 and :648-1021.    */
typedef struct { char name; int negated; int n; sre_char r[10]; } sre_esc_class_t;

/* outside brackets a negated escape is an NCLASS over the positive ranges   */
static const sre_esc_class_t  esc_positive[] = {
    { 'd', 0, 1, { '0', '9' } },
    { 'w', 0, 4, { 'A', 'Z', 'a', 'z', '0', '9', '_', '_' } },
    { 's', 0, 5, { ' ', ' ', '\f', '\f', '\n', '\n', '\r', '\r', '\t', '\t' } },
    { 'h', 0, 3, { 0x09, 0x09, 0x20, 0x20, 0xa0, 0xa0 } },
    { 'v', 0, 5, { 0x0a, 0x0a, 0x0b, 0x0b, 0x0c, 0x0c, 0x0d, 0x0d, 0x85, 0x85 } },
};

/* inside brackets a negated escape is spelled out as its complement         */
static const sre_esc_class_t  esc_complement[] = {
    { 'D', 1, 2, { 0, 47, 58, 255 } },
    { 'W', 1, 5, { 0, 47, 58, 64, 91, 94, 96, 96, 123, 255 } },
    { 'S', 1, 4, { 0, 8, 11, 11, 14, 31, 33, 255 } },
    { 'H', 1, 4, { 0x00, 0x08, 0x0a, 0x1f, 0x21, 0x9f, 0xa1, 0xff } },
    { 'V', 1, 3, { 0x00, 0x09, 0x0e, 0x84, 0x86, 0xff } },
};

static const sre_esc_class_t *
esc_lookup(const sre_esc_class_t *tab, size_t n, int name)
{
    size_t i;
    for (i = 0; i < n; i++) {
        if (tab[i].name == name) {
            return &tab[i];
        }
    }
    return NULL;
}

#define NELEMS(a)  (sizeof(a) / sizeof((a)[0]))

static int
is_print_c(int c)
{
    return c >= 0x20 && c <= 0x7e;  /* isprint() in the "C" locale */
}

static int
is_oct(int c)
{
    return c >= '0' && c <= '7';
}

static int
hex_val(int c)
{
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    return -1;
}

SRE_NOAPI sre_regex_t *
sre_regex_create(sre_pool_t *pool, sre_regex_type_t type, sre_regex_t *left,
    sre_regex_t *right)
{
    sre_regex_t *r = sre_pcalloc(pool, sizeof(sre_regex_t));
    if (r == NULL) {
        return NULL;
    }
    r->type = type;
    r->left = left;
    r->right = right;
    return r;
}

/* append [from,to] to a range list; returns the new tail (NULL on OOM) */
static sre_regex_range_t *
range_append(sre_parser_t *p, sre_regex_t *cls, sre_regex_range_t *tail,
    sre_char from, sre_char to)
{
    sre_regex_range_t *r = sre_palloc(p->pool, sizeof(sre_regex_range_t));
    if (r == NULL) {
        p->oom = 1;
        return NULL;
    }
    r->from = from;
    r->to = to;
    r->next = NULL;
    if (tail) {
        tail->next = r;
    } else {
        cls->data.range = r;
    }
    return r;
}

static sre_regex_t *
class_from_ranges(sre_parser_t *p, sre_regex_type_t type, const sre_char *r,
    int n)
{
    sre_regex_t        *cls = sre_regex_create(p->pool, type, NULL, NULL);
    sre_regex_range_t  *tail = NULL;
    int                 i;

    if (cls == NULL) {
        p->oom = 1;
        return NULL;
    }
    for (i = 0; i < n; i++) {
        tail = range_append(p, cls, tail, r[2 * i], r[2 * i + 1]);
        if (tail == NULL) {
            return NULL;
        }
    }
    return cls;
}

static sre_regex_t *
class_not_newline(sre_parser_t *p)
{
    static const sre_char nl[2] = { '\n', '\n' };
    return class_from_ranges(p, SRE_REGEX_TYPE_NCLASS, nl, 1);
}

/* ---- lexer --------------------------------------------------------------- */

static int
tok_set(sre_parser_t *p, int kind)
{
    p->tok.kind = kind;
    return kind;
}

static int
tok_char(sre_parser_t *p, unsigned c)
{
    p->tok.ch = (sre_char) c;
    return tok_set(p, TOK_CHAR);
}

static int
tok_node(sre_parser_t *p, int kind, sre_regex_t *re)
{
    if (re == NULL) {
        return tok_set(p, TOK_BAD);
    }
    p->tok.re = re;
    return tok_set(p, kind);
}

/*
 * \ddd, \o{...}, \x.. and \x{..} numeric escapes.  `in_class` selects the two
 * places where the bracket-expression dialect differs (.y:1135-1158 vs
 * :419-453, and :1261-1264).  Returns the byte value or -1 for a bad token.
 */
static int
lex_octal(sre_parser_t *p, int first, int in_class)
{
    unsigned  num = first - '0';
    int       ndigits = 1;

    while (ndigits < 3 && is_oct(*p->src)) {
        num = (num << 3) + (*p->src++ - '0');
        ndigits++;
    }
    if (ndigits == 3) {
        return num > 255 ? -1 : (int) num;
    }
    /* outside brackets a lone non-zero digit would be a back-reference */
    if (!in_class && ndigits == 1 && num != 0) {
        return -1;
    }
    return (int) num;
}

static int
lex_o_brace(sre_parser_t *p, int in_class)
{
    unsigned  num = 0;
    int       i = 0;
    sre_char  c;

    if (*p->src++ != '{') {
        return -1;
    }
    c = *p->src++;
    for (;;) {
        if (is_oct(c)) {
            num = (num << 3) + (c - '0');
        } else if (c == '}') {
            return (sre_char) num;
        } else if (in_class || c == '\0') {
            return -1;
        } else {
            p->src--;           /* not ours: hand it back */
            return (sre_char) num;
        }
        if (++i == 3) {
            if (*p->src++ != '}' || num > 255) {
                return -1;
            }
            return (sre_char) num;
        }
        c = *p->src++;
    }
}

static int
lex_hex(sre_parser_t *p, int in_class)
{
    unsigned  num = 0;
    int       i = 0, braced = 0, h;
    sre_char  c = *p->src++;

    if (c == '{') {
        braced = 1;
        c = *p->src++;
    }
    for (;;) {
        h = hex_val(c);
        if (h >= 0) {
            num = (num << 4) + h;
        } else if (braced) {
            return c == '}' ? (int) (sre_char) num : -1;
        } else if (in_class && c == '\0') {
            return -1;
        } else {
            p->src--;
            return (sre_char) num;
        }
        if (++i == 2) {
            if (braced && *p->src++ != '}') {
                return -1;
            }
            return (sre_char) num;
        }
        c = *p->src++;
    }
}

static int
simple_escape(int c)
{
    switch (c) {
    case 't': return '\t';
    case 'n': return '\n';
    case 'r': return '\r';
    case 'f': return '\f';
    case 'a': return 7;
    case 'e': return 27;
    default:  return -1;
    }
}

static int
lex_escape(sre_parser_t *p)
{
    const sre_esc_class_t  *ec;
    sre_regex_t            *r;
    int                     v;
    sre_char                c = *p->src++;

    if (c == '\0') {
        return tok_set(p, TOK_BAD);
    }
    if (!is_print_c(c) || strchr("'\" iM%@!,_-|*+?():.^$&\\/[]{}", c)) {
        return tok_char(p, c);
    }
    if (is_oct(c)) {
        v = lex_octal(p, c, 0);
        return v < 0 ? tok_set(p, TOK_BAD) : tok_char(p, v);
    }

    switch (c) {
    case 'c':
        c = *p->src++;
        if (c == '\0') {
            return tok_set(p, TOK_BAD);
        }
        if (c >= 'a' && c <= 'z') {
            c -= 32;
        }
        return tok_char(p, c ^ 64);

    case 'o':
        v = lex_o_brace(p, 0);
        return v < 0 ? tok_set(p, TOK_BAD) : tok_char(p, v);

    case 'x':
        v = lex_hex(p, 0);
        return v < 0 ? tok_set(p, TOK_BAD) : tok_char(p, v);

    case 'B': case 'b': case 'z': case 'A':
        r = sre_regex_create(p->pool, SRE_REGEX_TYPE_ASSERT, NULL, NULL);
        if (r) {
            r->data.assertion = c == 'B' ? SRE_REGEX_ASSERT_BIG_B
                              : c == 'b' ? SRE_REGEX_ASSERT_SMALL_B
                              : c == 'z' ? SRE_REGEX_ASSERT_SMALL_Z
                              : SRE_REGEX_ASSERT_BIG_A;
        }
        return tok_node(p, TOK_ASSERT, r);

    case 'N':
        return tok_node(p, TOK_CLASS, class_not_newline(p));

    case 'C':
        if (p->flags & SRE_REGEX_NEWLINE) {
            return tok_node(p, TOK_CLASS, class_not_newline(p));
        }
        return tok_node(p, TOK_CLASS,
                        sre_regex_create(p->pool, SRE_REGEX_TYPE_DOT, NULL,
                                         NULL));

    case '#':
        return tok_char(p, c);

    default:
        break;
    }

    v = simple_escape(c);
    if (v >= 0) {
        return tok_char(p, v);
    }

    /* \d \w \s \h \v and their upper-case negations */
    ec = esc_lookup(esc_positive, NELEMS(esc_positive), tolower(c));
    if (ec != NULL && c != 'n' && c != 'N') {
        return tok_node(p, TOK_CLASS,
                        class_from_ranges(p, c == ec->name
                                             ? SRE_REGEX_TYPE_CLASS
                                             : SRE_REGEX_TYPE_NCLASS,
                                          ec->r, ec->n));
    }

    return tok_set(p, TOK_BAD);
}

static int
lex_bracket(sre_parser_t *p)
{
    const sre_esc_class_t  *ec;
    sre_regex_t            *cls;
    sre_regex_range_t      *tail = NULL;
    int                     pending_dash = 0;   /* saw "x-", awaiting the end */
    int                     no_range = 0;       /* a '-' here is a literal    */
    int                     n = 0, v, i;
    sre_char                c;

    if (*p->src == '^') {
        p->src++;
        cls = sre_regex_create(p->pool, SRE_REGEX_TYPE_NCLASS, NULL, NULL);
    } else {
        cls = sre_regex_create(p->pool, SRE_REGEX_TYPE_CLASS, NULL, NULL);
    }
    if (cls == NULL) {
        return tok_set(p, TOK_BAD);
    }

    for (;;) {
        n++;
        c = *p->src++;

        if (c == '\0') {
            return tok_set(p, TOK_BAD);
        }

        if (c == ']' && n > 1) {
            if (pending_dash && range_append(p, cls, tail, '-', '-') == NULL) {
                return tok_set(p, TOK_BAD);
            }
            return tok_node(p, TOK_CLASS, cls);
        }

        if (c == '-' && !pending_dash && tail && !no_range) {
            pending_dash = 1;
            continue;
        }

        if (c == '\\') {
            c = *p->src++;
            if (c == '\0') {
                return tok_set(p, TOK_BAD);
            }
            if (is_oct(c)) {
                v = lex_octal(p, c, 1);
            } else if (c == 'c') {
                c = *p->src++;
                if (c == '\0') {
                    return tok_set(p, TOK_BAD);
                }
                if (c >= 'a' && c <= 'z') {
                    c -= 32;
                }
                v = (sre_char) (c ^ 64);
            } else if (c == 'o') {
                v = lex_o_brace(p, 1);
            } else if (c == 'x') {
                v = lex_hex(p, 1);
            } else if (c == 'b') {
                v = 8;
            } else if (simple_escape(c) >= 0) {
                v = simple_escape(c);
            } else if (c == '#' || !is_print_c(c)
                       || strchr("'\" iMzC%@!,_-|*+?():.^$&\\/[]{}", c))
            {
                v = c;
            } else {
                /* a class escape: never an end point of a range */
                ec = esc_lookup(esc_positive, NELEMS(esc_positive), c);
                if (ec == NULL) {
                    ec = esc_lookup(esc_complement, NELEMS(esc_complement), c);
                }
                if (ec == NULL) {
                    return tok_set(p, TOK_BAD);
                }
                if (pending_dash) {
                    tail = range_append(p, cls, tail, '-', '-');
                    if (tail == NULL) {
                        return tok_set(p, TOK_BAD);
                    }
                    pending_dash = 0;
                }
                for (i = 0; i < ec->n; i++) {
                    tail = range_append(p, cls, tail, ec->r[2 * i],
                                        ec->r[2 * i + 1]);
                    if (tail == NULL) {
                        return tok_set(p, TOK_BAD);
                    }
                }
                no_range = 1;
                continue;
            }
            if (v < 0) {
                return tok_set(p, TOK_BAD);
            }
            c = (sre_char) v;
        }

        /* an ordinary member byte */
        if (pending_dash) {
            tail->to = c;
            if (tail->to < tail->from) {
                return tok_set(p, TOK_BAD);
            }
            pending_dash = 0;
            no_range = 1;
            continue;
        }
        no_range = 0;
        tail = range_append(p, cls, tail, c, c);
        if (tail == NULL) {
            return tok_set(p, TOK_BAD);
        }
    }
}

/* "{n}", "{n,}", "{n,m}"; anything else leaves '{' a literal (.y:1693-1784) */
static int
lex_brace(sre_parser_t *p)
{
    const sre_char  *s = p->src;
    long             from = 0, to;

    if (!isdigit(*s)) {
        return tok_char(p, '{');
    }
    while (isdigit(*s)) {
        from = from * 10 + (*s++ - '0');
        if (from > 100000) {
            from = 100000;      /* saturate; >= 500 is rejected below */
        }
    }
    if (*s == '}') {
        to = from;
    } else if (*s == ',') {
        s++;
        if (*s == '}') {
            to = -1;
        } else if (isdigit(*s)) {
            to = 0;
            while (isdigit(*s)) {
                to = to * 10 + (*s++ - '0');
                if (to > 100000) {
                    to = 100000;
                }
            }
            if (*s != '}') {
                return tok_char(p, '{');
            }
        } else {
            return tok_char(p, '{');
        }
    } else {
        return tok_char(p, '{');
    }
    p->src = s + 1;

    if (from >= 500 || to >= 500 || (to >= 0 && from > to)) {
        return tok_set(p, TOK_BAD);
    }
    if (from == 0 && to == 1) {
        return tok_set(p, '?');
    }
    if (from == 0 && to == -1) {
        return tok_set(p, '*');
    }
    if (from == 1 && to == -1) {
        return tok_set(p, '+');
    }
    p->tok.from = (int) from;
    p->tok.to = (int) to;
    return tok_set(p, TOK_CQUANT);
}

static int
next_token(sre_parser_t *p)
{
    sre_char c;

    p->tok.pos = p->src;
    p->tok.re = NULL;

    if (p->src == NULL || *p->src == '\0') {
        return tok_set(p, TOK_EOF);
    }
    c = *p->src++;
    if (strchr("|*+?():.^$", c)) {
        return tok_set(p, c);
    }
    switch (c) {
    case '\\': return lex_escape(p);
    case '[':  return lex_bracket(p);
    case '{':  return lex_brace(p);
    default:   return tok_char(p, c);
    }
}

/* ---- parser -------------------------------------------------------------- */

static sre_regex_t *parse_alt(sre_parser_t *p);

static sre_regex_t *
syntax_error(sre_parser_t *p)
{
    if (p->err_pos == NULL) {
        p->err_pos = p->tok.pos;
    }
    return NULL;
}

static sre_regex_t *
node(sre_parser_t *p, sre_regex_type_t type, sre_regex_t *l, sre_regex_t *r)
{
    sre_regex_t *re = sre_regex_create(p->pool, type, l, r);
    if (re == NULL) {
        p->oom = 1;
    }
    return re;
}

static int
starts_atom(int kind)
{
    switch (kind) {
    case '(': case '.': case '^': case '$': case ':':
    case TOK_CHAR: case TOK_CLASS: case TOK_ASSERT:
        return 1;
    default:
        return 0;
    }
}

/* caseless: append the other-case image after each range that overlaps a
 * letter block, upper block first (behaviour of sre_regex.c:170-214) */
static int
fold_class(sre_parser_t *p, sre_regex_range_t *range)
{
    sre_regex_range_t  *r, *extra;
    sre_char            from, to;
    int                 pass;

    for (r = range; r; r = r->next) {
        from = r->from;
        to = r->to;
        for (pass = 0; pass < 2; pass++) {
            sre_char lo = pass == 0 ? 'A' : 'a';
            sre_char hi = pass == 0 ? 'Z' : 'z';
            int      delta = pass == 0 ? 32 : -32;

            if (to < lo || from > hi) {
                continue;
            }
            extra = sre_palloc(p->pool, sizeof(sre_regex_range_t));
            if (extra == NULL) {
                p->oom = 1;
                return SRE_ERROR;
            }
            extra->from = (from > lo ? from : lo) + delta;
            extra->to = (to < hi ? to : hi) + delta;
            extra->next = r->next;
            r->next = extra;
            r = extra;
        }
    }
    return SRE_OK;
}

static sre_regex_t *
parse_atom(sre_parser_t *p)
{
    sre_regex_t  *re, *body;
    sre_uint_t    group;
    sre_char      c;

    switch (p->tok.kind) {
    case '(':
        next_token(p);
        if (p->tok.kind == '?') {
            next_token(p);
            if (p->tok.kind != ':') {
                return syntax_error(p);
            }
            next_token(p);
            body = parse_alt(p);
            if (body == NULL) {
                return NULL;
            }
            if (p->tok.kind != ')') {
                return syntax_error(p);
            }
            next_token(p);
            return body;
        }
        group = ++(*p->ncaps);  /* numbered by opening parenthesis */
        body = parse_alt(p);
        if (body == NULL) {
            return NULL;
        }
        if (p->tok.kind != ')') {
            return syntax_error(p);
        }
        next_token(p);
        re = node(p, SRE_REGEX_TYPE_PAREN, body, NULL);
        if (re) {
            re->data.group = group;
        }
        return re;

    case TOK_CHAR:
        c = p->tok.ch;
        next_token(p);
        if ((p->flags & SRE_REGEX_CASELESS) && isalpha(c) && c < 0x80) {
            sre_char pair[4];
            pair[0] = pair[1] = c;
            pair[2] = pair[3] = c ^ 0x20;
            return class_from_ranges(p, SRE_REGEX_TYPE_CLASS, pair, 2);
        }
        re = node(p, SRE_REGEX_TYPE_LIT, NULL, NULL);
        if (re) {
            re->data.ch = c;
        }
        return re;

    case ':':
        next_token(p);
        re = node(p, SRE_REGEX_TYPE_LIT, NULL, NULL);
        if (re) {
            re->data.ch = ':';
        }
        return re;

    case '.':
        next_token(p);
        if (p->flags & SRE_REGEX_NEWLINE) {
            return class_not_newline(p);
        }
        return node(p, SRE_REGEX_TYPE_DOT, NULL, NULL);

    case '^':
    case '$':
        re = node(p, SRE_REGEX_TYPE_ASSERT, NULL, NULL);
        if (re) {
            re->data.assertion = p->tok.kind == '^' ? SRE_REGEX_ASSERT_CARET
                                                    : SRE_REGEX_ASSERT_DOLLAR;
        }
        next_token(p);
        return re;

    case TOK_ASSERT:
        re = p->tok.re;
        next_token(p);
        return re;

    case TOK_CLASS:
        re = p->tok.re;
        next_token(p);
        if ((p->flags & SRE_REGEX_CASELESS)
            && (re->type == SRE_REGEX_TYPE_CLASS
                || re->type == SRE_REGEX_TYPE_NCLASS)
            && fold_class(p, re->data.range) != SRE_OK)
        {
            return NULL;
        }
        return re;

    default:
        return syntax_error(p);
    }
}

/* x{n} = n copies; x{n,} = n copies + x*; x{n,m} = n copies + (m-n) x?.  The
 * copies are the *same* node, hence share capture-group numbers
 * (behaviour of .y:2011-2084). */
static sre_regex_t *
expand_counted(sre_parser_t *p, sre_regex_t *subj, int from, int to,
    unsigned greedy)
{
    sre_regex_t  *seq, *opt;
    int           i;

    if (from == 1 && to == 1) {
        return subj;
    }
    if (from == 0) {
        seq = node(p, SRE_REGEX_TYPE_NIL, NULL, NULL);
    } else {
        seq = subj;
        for (i = 1; seq && i < from; i++) {
            seq = node(p, SRE_REGEX_TYPE_CAT, seq, subj);
        }
    }
    if (seq == NULL || from == to) {
        return seq;
    }
    opt = node(p, to < 0 ? SRE_REGEX_TYPE_STAR : SRE_REGEX_TYPE_QUEST, subj,
               NULL);
    if (opt == NULL) {
        return NULL;
    }
    opt->data.greedy = greedy;
    if (to < 0) {
        return node(p, SRE_REGEX_TYPE_CAT, seq, opt);
    }
    for (i = from; seq && i < to; i++) {
        seq = node(p, SRE_REGEX_TYPE_CAT, seq, opt);
    }
    return seq;
}

static sre_regex_t *
parse_repeat(sre_parser_t *p)
{
    sre_regex_t  *atom, *re;
    int           kind, from, to;
    unsigned      greedy = 1;

    atom = parse_atom(p);
    if (atom == NULL) {
        return NULL;
    }
    kind = p->tok.kind;
    if (kind != '*' && kind != '+' && kind != '?' && kind != TOK_CQUANT) {
        return atom;
    }
    from = p->tok.from;
    to = p->tok.to;
    next_token(p);
    if (p->tok.kind == '?') {
        greedy = 0;
        next_token(p);
    }
    if (kind == TOK_CQUANT) {
        return expand_counted(p, atom, from, to, greedy);
    }
    re = node(p, kind == '*' ? SRE_REGEX_TYPE_STAR
                 : kind == '+' ? SRE_REGEX_TYPE_PLUS : SRE_REGEX_TYPE_QUEST,
              atom, NULL);
    if (re) {
        re->data.greedy = greedy;
    }
    return re;
}

static sre_regex_t *
parse_concat(sre_parser_t *p)
{
    sre_regex_t *seq, *next;

    if (!starts_atom(p->tok.kind)) {
        return node(p, SRE_REGEX_TYPE_NIL, NULL, NULL);
    }
    seq = parse_repeat(p);
    while (seq && starts_atom(p->tok.kind)) {
        next = parse_repeat(p);
        if (next == NULL) {
            return NULL;
        }
        seq = node(p, SRE_REGEX_TYPE_CAT, seq, next);
    }
    return seq;
}

static sre_regex_t *
parse_alt(sre_parser_t *p)
{
    sre_regex_t *alt = parse_concat(p), *rhs;

    while (alt && p->tok.kind == '|') {
        next_token(p);
        rhs = parse_concat(p);
        if (rhs == NULL) {
            return NULL;
        }
        alt = node(p, SRE_REGEX_TYPE_ALT, alt, rhs);
    }
    return alt;
}

/* one user regex, up to the end of the string */
static sre_regex_t *
parse_one(sre_pool_t *pool, const sre_char *src, sre_uint_t *ncaps, int flags,
    const sre_char **err_pos)
{
    sre_parser_t  p;
    sre_regex_t  *re;

    memset(&p, 0, sizeof(p));
    p.pool = pool;
    p.src = src;
    p.flags = flags;
    p.ncaps = ncaps;

    next_token(&p);
    re = parse_alt(&p);
    if (re != NULL && p.tok.kind != TOK_EOF) {
        syntax_error(&p);
        re = NULL;
    }
    if (re == NULL && !p.oom) {
        *err_pos = p.err_pos;
    }
    return re;
}

/* TOPLEVEL(id, Paren(group, user)) */
static sre_regex_t *
wrap_toplevel(sre_pool_t *pool, sre_regex_t *user, sre_uint_t group,
    sre_int_t regex_id)
{
    sre_regex_t *re = sre_regex_create(pool, SRE_REGEX_TYPE_PAREN, user, NULL);
    if (re == NULL) {
        return NULL;
    }
    re->data.group = group;
    re = sre_regex_create(pool, SRE_REGEX_TYPE_TOPLEVEL, re, NULL);
    if (re) {
        re->data.regex_id = regex_id;
    }
    return re;
}

/* Cat(NgStar(Dot), body): the unanchored-search prefix ".*?" (.y:1830-1857);
 * its Dot matches newlines even under SRE_REGEX_NEWLINE */
static sre_regex_t *
wrap_search(sre_pool_t *pool, sre_regex_t *body, sre_uint_t nregexes,
    sre_uint_t *multi_ncaps)
{
    sre_regex_t *dot, *star, *re;

    dot = sre_regex_create(pool, SRE_REGEX_TYPE_DOT, NULL, NULL);
    if (dot == NULL) {
        return NULL;
    }
    star = sre_regex_create(pool, SRE_REGEX_TYPE_STAR, dot, NULL);
    if (star == NULL) {
        return NULL;
    }
    re = sre_regex_create(pool, SRE_REGEX_TYPE_CAT, star, body);
    if (re == NULL) {
        return NULL;
    }
    re->nregexes = nregexes;
    re->data.multi_ncaps = multi_ncaps;
    return re;
}

SRE_API sre_regex_t *
sre_regex_parse(sre_pool_t *pool, sre_char *src, sre_uint_t *ncaps, int flags,
    sre_int_t *err_offset)
{
    const sre_char  *err_pos = NULL;
    sre_regex_t     *re;
    sre_uint_t      *multi_ncaps;

    *ncaps = 0;
    *err_offset = -1;

    re = parse_one(pool, src, ncaps, flags, &err_pos);
    if (re == NULL) {
        if (err_pos) {
            *err_offset = (sre_int_t) (err_pos - src);
        }
        return NULL;
    }
    re = wrap_toplevel(pool, re, 0, 0);
    if (re == NULL) {
        return NULL;
    }
    multi_ncaps = sre_palloc(pool, sizeof(sre_uint_t));
    if (multi_ncaps == NULL) {
        return NULL;
    }
    multi_ncaps[0] = *ncaps;
    return wrap_search(pool, re, 1, multi_ncaps);
}

SRE_API sre_regex_t *
sre_regex_parse_multi(sre_pool_t *pool, sre_char **regexes, sre_int_t nregexes,
    sre_uint_t *max_ncaps, int *multi_flags, sre_int_t *err_offset,
    sre_int_t *err_regex_id)
{
    const sre_char  *err_pos = NULL;
    sre_regex_t     *re, *all = NULL;
    sre_uint_t      *multi_ncaps, ncaps = 0, base;
    sre_int_t        i;

    *max_ncaps = 0;
    *err_offset = -1;
    *err_regex_id = -1;

    if (nregexes <= 0) {
        return NULL;
    }
    multi_ncaps = sre_palloc(pool, nregexes * sizeof(sre_uint_t));
    if (multi_ncaps == NULL) {
        return NULL;
    }

    /* re0|re1|...: group numbers run on across the set (.y:1902-1962) */
    for (i = 0; i < nregexes; i++) {
        *err_regex_id = i;
        base = ncaps;
        re = parse_one(pool, regexes[i], &ncaps,
                       multi_flags ? multi_flags[i] : 0, &err_pos);
        if (re == NULL) {
            if (err_pos) {
                *err_offset = (sre_int_t) (err_pos - regexes[i]);
            }
            return NULL;
        }
        re = wrap_toplevel(pool, re, base, i);
        if (re == NULL) {
            return NULL;
        }
        multi_ncaps[i] = ncaps - base;
        if (multi_ncaps[i] > *max_ncaps) {
            *max_ncaps = multi_ncaps[i];
        }
        ncaps++;    /* this regex's own $0 */
        if (all == NULL) {
            all = re;
        } else {
            all = sre_regex_create(pool, SRE_REGEX_TYPE_ALT, all, re);
            if (all == NULL) {
                return NULL;
            }
        }
    }
    return wrap_search(pool, all, nregexes, multi_ncaps);
}

/* ---- AST dump (format of sre_regex.c:36-166) ----------------------------- */

static void
dump_ranges(const char *name, sre_regex_range_t *r)
{
    printf("%s(", name);
    for (; r; r = r->next) {
        printf("[%d, %d]", r->from, r->to);
    }
    printf(")");
}

SRE_API void
sre_regex_dump(sre_regex_t *r)
{
    static const char *rep[] = { "Quest", "Star", "Plus" };
    const char        *s;

    switch (r->type) {
    case SRE_REGEX_TYPE_ALT:
    case SRE_REGEX_TYPE_CAT:
        printf(r->type == SRE_REGEX_TYPE_ALT ? "Alt(" : "Cat(");
        sre_regex_dump(r->left);
        printf(", ");
        sre_regex_dump(r->right);
        printf(")");
        break;
    case SRE_REGEX_TYPE_LIT:
        printf("Lit(%d)", (int) r->data.ch);
        break;
    case SRE_REGEX_TYPE_DOT:
        printf("Dot");
        break;
    case SRE_REGEX_TYPE_PAREN:
        printf("Paren(%lu, ", (unsigned long) r->data.group);
        sre_regex_dump(r->left);
        printf(")");
        break;
    case SRE_REGEX_TYPE_QUEST:
    case SRE_REGEX_TYPE_STAR:
    case SRE_REGEX_TYPE_PLUS:
        printf("%s%s(", r->data.greedy ? "" : "Ng",
               rep[r->type - SRE_REGEX_TYPE_QUEST]);
        sre_regex_dump(r->left);
        printf(")");
        break;
    case SRE_REGEX_TYPE_NIL:
        printf("Nil");
        break;
    case SRE_REGEX_TYPE_CLASS:
        dump_ranges("CLASS", r->data.range);
        break;
    case SRE_REGEX_TYPE_NCLASS:
        dump_ranges("NCLASS", r->data.range);
        break;
    case SRE_REGEX_TYPE_ASSERT:
        switch (r->data.assertion) {
        case SRE_REGEX_ASSERT_BIG_A:   s = "\\A"; break;
        case SRE_REGEX_ASSERT_CARET:   s = "^";   break;
        case SRE_REGEX_ASSERT_DOLLAR:  s = "$";   break;
        case SRE_REGEX_ASSERT_SMALL_Z: s = "\\z"; break;
        case SRE_REGEX_ASSERT_BIG_B:   s = "\\B"; break;
        case SRE_REGEX_ASSERT_SMALL_B: s = "\\b"; break;
        default:                       s = "???"; break;
        }
        printf("ASSERT(%s)", s);
        break;
    case SRE_REGEX_TYPE_TOPLEVEL:
        printf("TOPLEVEL(%lu, ", (unsigned long) r->data.regex_id);
        sre_regex_dump(r->left);
        printf(")");
        break;
    default:
        printf("???");
        break;
    }
}
