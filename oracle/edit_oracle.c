/*
 * edit_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle) for the "next" row N3 of SURVEY.md 8f:
 * the pairwise distance amplicon_sorter computes with edlib,
 *
 *   /root/reference/scripts/auxiliary_code/amplicon_sorter.py:225-235   distance(X1, X2, mode='NW')
 *   /root/reference/scripts/auxiliary_code/amplicon_sorter.py:838-849   distance_finetune()  (mode='HW')
 *       s = edlib.align(A1, A2, task='distance', mode=mode)      # A1 = the shorter sequence
 *       iden = round(1 - s['editDistance'] / len(A2), 3)
 *
 * edlib (un-vendored C++ dependency of the reference) returns the exact unit-cost edit distance:
 * NW = global (Levenshtein) distance, HW = the query aligned to the best infix of the target (gaps
 * before and after the query in the target are free).  Characters are compared by exact byte
 * equality (edlib's default: no additional equalities).  Unlike the cutadapt path this oracle IS
 * pinned: the distance is defined by the textbook recurrence below, there is no implementation
 * choice to restate.
 */
#include <stdint.h>
#include <stdlib.h>
#include <pthread.h>

/* mode 0 = NW, 1 = HW.  q = query (length m), t = target (length n). */
uint32_t oracle_edit_distance(const uint8_t *q, uint32_t m, const uint8_t *t, uint32_t n, int mode)
{
    uint32_t *col = (uint32_t *)malloc(((size_t)m + 1) * sizeof(uint32_t));
    for (uint32_t i = 0; i <= m; i++) col[i] = i;                  /* D[i][0] = i */
    uint32_t best = m;                                               /* HW: min over the end columns */
    for (uint32_t j = 1; j <= n; j++) {
        uint32_t diag = col[0];
        col[0] = mode ? 0u : j;                                      /* D[0][j] = j (NW) or 0 (HW) */
        for (uint32_t i = 1; i <= m; i++) {
            uint32_t c = diag + (q[i - 1] != t[j - 1]);
            if (col[i] + 1 < c) c = col[i] + 1;
            if (col[i - 1] + 1 < c) c = col[i - 1] + 1;
            diag = col[i];
            col[i] = c;
        }
        if (col[m] < best) best = col[m];
    }
    uint32_t r = mode ? best : col[m];
    free(col);
    return r;
}

typedef struct {
    const uint8_t *seqs; const uint64_t *off; const uint32_t *len;
    const uint32_t *pa, *pb; uint64_t lo, hi; int mode; uint32_t *out;
} job_t;

static void *worker(void *p)
{
    job_t *j = (job_t *)p;
    for (uint64_t k = j->lo; k < j->hi; k++) {
        uint32_t a = j->pa[k], b = j->pb[k];
        /* amplicon_sorter: the longer sequence is the target; X1 stays the query on equal lengths */
        if (j->len[a] > j->len[b]) { uint32_t x = a; a = b; b = x; }
        j->out[k] = oracle_edit_distance(j->seqs + j->off[a], j->len[a], j->seqs + j->off[b], j->len[b], j->mode);
    }
    return 0;
}

void oracle_edit_distances(const uint8_t *seqs, const uint64_t *off, const uint32_t *len,
                           const uint32_t *pa, const uint32_t *pb, uint64_t n_pairs, int mode,
                           uint32_t *out, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    job_t jobs[256];
    for (int i = 0; i < n_threads; i++) {
        jobs[i] = (job_t){seqs, off, len, pa, pb, n_pairs * (uint64_t)i / n_threads,
                          n_pairs * (uint64_t)(i + 1) / n_threads, mode, out};
        pthread_create(&th[i], 0, worker, &jobs[i]);
    }
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], 0);
}
