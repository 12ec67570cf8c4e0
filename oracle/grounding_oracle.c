/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the rule-grounding hot path.
 *
 * Plain-C restatement of the reference algorithm
 *   /root/reference/src/data.py:136-173  (KnowledgeGraph.grounding / propagate)
 *   /root/reference/src/trainer.py:189-201 (filtered rank bounds L, H)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path never does.
 *
 * Layout matches the reference: the frontier is x[N][B] int64 (entity-major), the
 * adjacency of one relation is two int64 arrays in train.txt order
 * (node_in = head of edge k, node_out = tail of edge k; data.py:63-64,151-152).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* One hop, data.py:149-173.
 *   message[k][b] = x[node_in[k]][b]                         (gather,  :157)
 *   message[etr[b]][b] = 0  when etr != NULL                  (masking, :164-170)
 *   out[node_out[k]][b] += message[k][b]                      (scatter-sum, :161/:171)
 * `out` must hold N*B int64 and is overwritten. */
void oracle_propagate(int64_t N, int64_t B, int64_t E,
                      const int64_t *node_in, const int64_t *node_out,
                      const int64_t *x, const int64_t *etr, int64_t *out)
{
    memset(out, 0, (size_t)(N * B) * sizeof(int64_t));
    for (int64_t k = 0; k < E; ++k) {
        const int64_t *src = x + node_in[k] * B;
        int64_t *dst = out + node_out[k] * B;
        if (etr) {
            for (int64_t b = 0; b < B; ++b)
                dst[b] += (etr[b] == k) ? 0 : src[b];
        } else {
            for (int64_t b = 0; b < B; ++b)
                dst[b] += src[b];
        }
    }
}

/* Full rule body, data.py:136-147.  rel_ptr[R+1] indexes node_in/node_out which hold the
 * per-relation edge lists back to back (each in train.txt order).  The query edge is
 * removed only on hops whose relation equals the head relation r (data.py:143-146).
 * Output is counts[B][N] int64 (the reference returns x.squeeze(-1).transpose(0,1)). */
int oracle_grounding(int64_t N, int64_t B, const int64_t *rel_ptr,
                     const int64_t *node_in, const int64_t *node_out,
                     const int64_t *h, int64_t r,
                     const int64_t *body, int64_t L,
                     const int64_t *etr, int64_t *counts)
{
    int64_t *x = (int64_t *)calloc((size_t)(N * B), sizeof(int64_t));
    int64_t *y = (int64_t *)malloc((size_t)(N * B) * sizeof(int64_t));
    if (!x || !y) { free(x); free(y); return -1; }
    for (int64_t b = 0; b < B; ++b) x[h[b] * B + b] = 1;       /* one_hot, :139 */
    for (int64_t i = 0; i < L; ++i) {
        int64_t rho = body[i];
        int64_t e0 = rel_ptr[rho], e1 = rel_ptr[rho + 1];
        oracle_propagate(N, B, e1 - e0, node_in + e0, node_out + e0, x,
                         (rho == r) ? etr : NULL, y);
        int64_t *t = x; x = y; y = t;
    }
    for (int64_t e = 0; e < N; ++e)
        for (int64_t b = 0; b < B; ++b)
            counts[b * N + e] = x[e * B + b];
    free(x); free(y);
    return 0;
}

/* trainer.py:189-201.  logits fp32[Q][N], flag u8[Q][N] (1 = entity takes part in the
 * ranking), mask u8[Q][N], t[Q].  LH[Q][2] = (L, H). */
void oracle_filtered_rank(int64_t Q, int64_t N, const float *logits,
                          const uint8_t *flag, const uint8_t *mask,
                          const int64_t *t, int64_t *LH)
{
    for (int64_t k = 0; k < Q; ++k) {
        const float *row = logits + k * N;
        if (mask[k * N + t[k]]) {
            float val = row[t[k]];
            int64_t gt = 0, ge = 0;
            for (int64_t e = 0; e < N; ++e) {
                if (!flag[k * N + e]) continue;
                gt += row[e] > val;
                ge += row[e] >= val;
            }
            LH[2 * k] = gt + 1;
            LH[2 * k + 1] = ge + 2;
        } else {
            LH[2 * k] = 1;
            LH[2 * k + 1] = N + 1;
        }
    }
}
