/*
 * sre_pool.c -- arena behind sre_create_pool / sre_reset_pool /
 * sre_destroy_pool (API of reference sregex.h:82-84; the reference's
 * implementation is the nginx pool in sre_palloc.c).  Ours is a plain chain of
 * malloc'ed slabs with bump allocation plus a cleanup list, which the CUDA side
 * uses to release device buffers that belong to pool-allocated contexts.
 * Unlike the reference (sre_palloc.c:118-148, which never rewinds
 * pool->current) a reset really returns the pool to its initial state.
 */
#include "sre_internal.h"

typedef struct sre_slab_s  sre_slab_t;
struct sre_slab_s {
    sre_slab_t  *next;
    size_t       cap, used;
    /* payload follows, 16-byte aligned */
};

typedef struct sre_cleanup_s  sre_cleanup_t;
struct sre_cleanup_s {
    sre_pool_cleanup_pt  handler;
    void                *data;
    sre_cleanup_t       *next;
};

struct sre_pool_s {
    size_t          slab_size;
    sre_slab_t     *slabs;      /* newest first */
    sre_cleanup_t  *cleanups;
};

#define SRE_SLAB_HDR  ((sizeof(sre_slab_t) + 15) & ~(size_t) 15)

static sre_slab_t *
sre_slab_new(size_t cap)
{
    sre_slab_t *s = malloc(SRE_SLAB_HDR + cap);
    if (s == NULL) {
        return NULL;
    }
    s->next = NULL;
    s->cap = cap;
    s->used = 0;
    return s;
}

SRE_API sre_pool_t *
sre_create_pool(size_t size)
{
    sre_pool_t *pool = malloc(sizeof(sre_pool_t));
    if (pool == NULL) {
        return NULL;
    }
    if (size < 256) {
        size = 256;
    }
    pool->slab_size = size;
    pool->slabs = NULL;
    pool->cleanups = NULL;
    return pool;
}

static void
sre_pool_release(sre_pool_t *pool)
{
    sre_cleanup_t  *c;
    sre_slab_t     *s, *n;

    /* cleanup records live inside the slabs: run them before freeing */
    for (c = pool->cleanups; c; c = c->next) {
        if (c->handler) {
            c->handler(c->data);
        }
    }
    pool->cleanups = NULL;
    for (s = pool->slabs; s; s = n) {
        n = s->next;
        free(s);
    }
    pool->slabs = NULL;
}

SRE_API void
sre_reset_pool(sre_pool_t *pool)
{
    if (pool) {
        sre_pool_release(pool);
    }
}

SRE_API void
sre_destroy_pool(sre_pool_t *pool)
{
    if (pool) {
        sre_pool_release(pool);
        free(pool);
    }
}

SRE_NOAPI void *
sre_palloc(sre_pool_t *pool, size_t size)
{
    sre_slab_t  *s = pool->slabs;
    size_t       need = (size + 15) & ~(size_t) 15, cap;
    void        *p;

    if (s == NULL || s->cap - s->used < need) {
        cap = need > pool->slab_size ? need : pool->slab_size;
        s = sre_slab_new(cap);
        if (s == NULL) {
            return NULL;
        }
        /* keep a partially filled slab at the head when the request was an
         * oversized one, so small allocations continue to fill it */
        if (pool->slabs && need > pool->slab_size) {
            s->next = pool->slabs->next;
            pool->slabs->next = s;
        } else {
            s->next = pool->slabs;
            pool->slabs = s;
        }
    }
    p = (char *) s + SRE_SLAB_HDR + s->used;
    s->used += need;
    return p;
}

SRE_NOAPI void *
sre_pcalloc(sre_pool_t *pool, size_t size)
{
    void *p = sre_palloc(pool, size);
    if (p) {
        memset(p, 0, size);
    }
    return p;
}

SRE_NOAPI int
sre_pool_add_cleanup(sre_pool_t *pool, sre_pool_cleanup_pt handler, void *data)
{
    sre_cleanup_t *c = sre_palloc(pool, sizeof(sre_cleanup_t));
    if (c == NULL) {
        return SRE_ERROR;
    }
    c->handler = handler;
    c->data = data;
    c->next = pool->cleanups;
    pool->cleanups = c;
    return SRE_OK;
}
