/*
 * workloads.c — CPU ORACLE / CPU BASELINE for the fixed-width benchmark shapes. TEST INFRASTRUCTURE ONLY
 * (same rule as tplx_oracle.c: never linked or called by the product path).
 *
 * Hand-lowered stage functions, written the way the reference's generated code runs them
 * (per partition: rows in order, filters short-circuit, aggregate folded sequentially from the initial
 * value, partition partials combined in order — tuplex/core/src/physical/TuplexSourceTaskBuilder.cc:104-215,
 * PipelineBuilder.cc:2525-2608, TransformTask.cc:218-299), run partition-parallel over host threads like
 * LocalBackend::performTasks (tuplex/core/src/ee/local/LocalBackend.cc:1531-1586).
 *
 *   Q6:  benchmarks/tpch/Q06/runtuplex.py:96-99
 *        filter(19940101 <= shipdate < 19950101).filter(0.05 <= discount <= 0.07).filter(quantity < 24)
 *        .aggregate(lambda a, b: a + b, lambda a, x: a + x[1] * x[2], 0.0)
 *   C1:  BASELINE.json configs[0]: map(x*x).filter(x % 2 == 0)
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int64_t tplx_o_floormod(int64_t x, int64_t y);

/* one task = one partition of rows [lo, hi) */
static double q6_partition(const int64_t *qty, const double *price, const double *disc, const int64_t *ship, uint64_t lo, uint64_t hi) {
    double a = 0.0; /* intermediate initialised from the initial value (BlockBasedTaskBuilder.cc:185-206) */
    for (uint64_t i = lo; i < hi; ++i) {
        int64_t sd = ship[i];
        if (!(19940101 <= sd && sd < 19950101)) continue;
        double d = disc[i];
        if (!(0.05 <= d && d <= 0.07)) continue;
        if (!(qty[i] < 24)) continue;
        a = a + price[i] * d;
    }
    return a;
}

typedef struct q6_job {
    const int64_t *qty;
    const double *price, *disc;
    const int64_t *ship;
    uint64_t n, part_rows;
    uint64_t n_parts;
    double *partials;
    volatile uint64_t *next;
} q6_job;

static void *q6_worker(void *arg) {
    q6_job *j = (q6_job *)arg;
    for (;;) {
        uint64_t p = __sync_fetch_and_add(j->next, 1);
        if (p >= j->n_parts) break;
        uint64_t lo = p * j->part_rows, hi = lo + j->part_rows;
        if (hi > j->n) hi = j->n;
        j->partials[p] = q6_partition(j->qty, j->price, j->disc, j->ship, lo, hi);
    }
    return NULL;
}

/* Returns the aggregate; partition partials are combined in partition order starting from 0.0
 * (deterministic stand-in for the thread-slot order of TransformTask.cc:278-299). */
double tplx_oracle_q6(const int64_t *qty, const double *price, const double *disc, const int64_t *ship, uint64_t n,
                      uint64_t part_rows, int threads) {
    if (part_rows == 0) part_rows = n ? n : 1;
    q6_job j;
    j.qty = qty; j.price = price; j.disc = disc; j.ship = ship; j.n = n; j.part_rows = part_rows;
    j.n_parts = (n + part_rows - 1) / part_rows;
    j.partials = (double *)calloc(j.n_parts ? j.n_parts : 1, sizeof(double));
    volatile uint64_t next = 0;
    j.next = &next;
    if (threads <= 1) q6_worker(&j);
    else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, q6_worker, &j);
        for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    double a = 0.0;
    for (uint64_t p = 0; p < j.n_parts; ++p) a = a + j.partials[p];
    free(j.partials);
    return a;
}

typedef struct c1_job {
    const int64_t *x;
    int64_t *out;          /* per partition output region = same offsets as the input */
    uint64_t *counts;
    uint64_t n, part_rows, n_parts;
    volatile uint64_t *next;
} c1_job;

static void *c1_worker(void *arg) {
    c1_job *j = (c1_job *)arg;
    for (;;) {
        uint64_t p = __sync_fetch_and_add(j->next, 1);
        if (p >= j->n_parts) break;
        uint64_t lo = p * j->part_rows, hi = lo + j->part_rows, k = lo;
        if (hi > j->n) hi = j->n;
        for (uint64_t i = lo; i < hi; ++i) {
            int64_t y = (int64_t)((uint64_t)j->x[i] * (uint64_t)j->x[i]); /* wrapping mul */
            if (tplx_o_floormod(y, 2) == 0) j->out[k++] = y;
        }
        j->counts[p] = k - lo;
    }
    return NULL;
}

/* map(x*x).filter(x%2==0); output compacted in order into out (capacity n); returns row count */
uint64_t tplx_oracle_c1(const int64_t *x, uint64_t n, int64_t *out, uint64_t part_rows, int threads) {
    if (part_rows == 0) part_rows = n ? n : 1;
    c1_job j;
    j.x = x; j.n = n; j.part_rows = part_rows; j.n_parts = (n + part_rows - 1) / part_rows;
    j.out = out;
    j.counts = (uint64_t *)calloc(j.n_parts ? j.n_parts : 1, 8);
    volatile uint64_t next = 0;
    j.next = &next;
    if (threads <= 1) c1_worker(&j);
    else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, c1_worker, &j);
        for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    /* concatenate partition outputs in partition order (LocalBackend.cc:1104-1152) */
    uint64_t w = 0;
    for (uint64_t p = 0; p < j.n_parts; ++p) {
        uint64_t lo = p * part_rows;
        if (w != lo) memmove(out + w, out + lo, j.counts[p] * 8);
        w += j.counts[p];
    }
    free(j.counts);
    return w;
}
