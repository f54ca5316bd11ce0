/*
 * join_oracle.c — CPU ORACLE of the hash join. TEST INFRASTRUCTURE ONLY (same rules as tplx_oracle.c).
 *
 * Row-at-a-time restatement of the reference's hash join (paths relative to /root/reference/tuplex/):
 *   build side   TransformTask::writeRowToHashTable (core/src/physical/TransformTask.cc:769-789): every build row is appended
 *                to the bucket of its key (extend_bucket: rows stay in insertion order); a NULL key (key == nullptr) goes to
 *                the null bucket. Tasks are merged in task order, so bucket order = input order of the build side.
 *   probe side   PipelineBuilder::addHashJoinProbe (core/src/physical/PipelineBuilder.cc:2330-2523): hashmap_get on the probe
 *                row's key (NULL key -> null bucket), then createInnerJoinBucketLoop (:2110-2212) emits one row per bucket row
 *                in bucket order; createLeftJoinBucketLoop (:2214-2328) runs the loop at least once with match_found = false,
 *                i.e. a probe row without a match is emitted once with NULL build columns.
 *   result shape JoinOperator::inferSchema (core/src/logical/JoinOperator.cc:163-184): left non-key columns, key, right non-key
 *                columns — assembled by the test helper from the index pairs this file returns.
 * Strings are compared as the reference's hashmap compares them: length + bytes of the NUL-terminated key.
 * Deliberately independent of the CUDA side: chained buckets with malloc'd row lists, not an open-addressing CSR.
 *
 * Parity pinned by the reference's own goldens: tuplex/test/core/JoinTest.cc:21-460 (tests/test_join.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct jcol {
    uint8_t type; /* tplx_type: 0 i64, 2 bool, 3 str */
    uint8_t pad[7];
    const void *data;
    const uint32_t *offsets;
    const uint32_t *valid; /* bit set = value present; NULL = all present */
} jcol;

typedef struct jbucket {
    struct jbucket *next;
    uint64_t hash;
    const char *key;
    uint64_t key_len;
    uint64_t *rows;
    uint64_t n, cap;
} jbucket;

static int j_isnull(const jcol *c, uint64_t r) { return c->valid && !((c->valid[r >> 5] >> (r & 31)) & 1u); }

static void j_key(const jcol *c, uint64_t r, const char **p, uint64_t *len) {
    if (c->type == 3) {
        *p = (const char *)c->data + c->offsets[r];
        *len = c->offsets[r + 1] - c->offsets[r];
    } else {
        *p = (const char *)c->data + 8 * r;
        *len = 8;
    }
}

static uint64_t j_hash(const char *p, uint64_t len) {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < len; ++i) h = (h ^ (uint8_t)p[i]) * 1099511628211ull;
    return h;
}

static void j_append(jbucket *b, uint64_t row) {
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 2;
        b->rows = (uint64_t *)realloc(b->rows, b->cap * 8);
    }
    b->rows[b->n++] = row;
}

/* Returns the number of output rows. out_probe / out_build (may be NULL: count only) receive, per output row, the probe row and
 * the build row (-1: left join without a match). */
uint64_t tplx_oracle_join(const jcol *build, uint64_t n_build, const jcol *probe, uint64_t n_probe, int left_outer, int64_t *out_probe,
                          int64_t *out_build, uint64_t cap) {
    uint64_t nb = 64;
    while (nb < 2 * n_build) nb <<= 1;
    jbucket **tab = (jbucket **)calloc(nb, sizeof(jbucket *));
    jbucket nullb;
    memset(&nullb, 0, sizeof(nullb));
    for (uint64_t r = 0; r < n_build; ++r) {
        if (j_isnull(build, r)) {
            j_append(&nullb, r);
            continue;
        }
        const char *k;
        uint64_t kl;
        j_key(build, r, &k, &kl);
        const uint64_t h = j_hash(k, kl);
        jbucket *b = tab[h & (nb - 1)];
        while (b && !(b->hash == h && b->key_len == kl && memcmp(b->key, k, kl) == 0)) b = b->next;
        if (!b) {
            b = (jbucket *)calloc(1, sizeof(jbucket));
            b->hash = h;
            b->key = k;
            b->key_len = kl;
            b->next = tab[h & (nb - 1)];
            tab[h & (nb - 1)] = b;
        }
        j_append(b, r);
    }
    uint64_t n_out = 0;
    for (uint64_t r = 0; r < n_probe; ++r) {
        const jbucket *b = NULL;
        if (j_isnull(probe, r)) {
            b = nullb.n ? &nullb : NULL;
        } else {
            const char *k;
            uint64_t kl;
            j_key(probe, r, &k, &kl);
            const uint64_t h = j_hash(k, kl);
            b = tab[h & (nb - 1)];
            while (b && !(b->hash == h && b->key_len == kl && memcmp(b->key, k, kl) == 0)) b = b->next;
        }
        if (b) {
            for (uint64_t j = 0; j < b->n; ++j) {
                if (out_probe && n_out < cap) {
                    out_probe[n_out] = (int64_t)r;
                    out_build[n_out] = (int64_t)b->rows[j];
                }
                ++n_out;
            }
        } else if (left_outer) {
            if (out_probe && n_out < cap) {
                out_probe[n_out] = (int64_t)r;
                out_build[n_out] = -1;
            }
            ++n_out;
        }
    }
    for (uint64_t i = 0; i < nb; ++i) {
        jbucket *b = tab[i];
        while (b) {
            jbucket *nx = b->next;
            free(b->rows);
            free(b);
            b = nx;
        }
    }
    free(nullb.rows);
    free(tab);
    return n_out;
}
