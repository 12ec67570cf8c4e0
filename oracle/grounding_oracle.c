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

/* ------------------------------------------------------------------------------------------------
 * Rule discovery, restating miner/rnnlogic.cpp:350-382 (KnowledgeGraph::rule_search) and :505-589
 * (RuleMiner::search_thread): for every triple (h, r, t) a recursive depth-first search from h with the
 * triple removed; a path that reaches t yields the rule r <- path and stops there; h == t yields the
 * empty body; r <- r is dropped.  The union over the triples is returned as sorted distinct 64-bit keys
 *     head << (b*L+3) | length << (b*L) | body[i] << (b*(L-1-i))      (b = rel_bits, L = max_len)
 * whose order is the order of the reference's rule list: head relation, then std::set<Rule> order (rnnlogic.cpp:118-133, 575-585).
 * adj_ptr[N+1] / adj_rel / adj_dst: out-edges by source entity (the miner's e2r2n).
 * Returns the number of distinct rules (keys beyond cap are counted but not written). */
typedef struct {
    const int32_t *adj_ptr, *adj_rel, *adj_dst;
    int max_len, rel_bits;
    uint64_t *set;           /* open addressing, ~0 = empty */
    uint64_t mask, used;
} miner_t;

static void miner_put(miner_t *m, uint64_t key)
{
    uint64_t x = key;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    uint64_t s = x & m->mask;
    while (m->set[s] != ~0ULL) {
        if (m->set[s] == key) return;
        s = (s + 1) & m->mask;
    }
    m->set[s] = key;
    m->used++;
    if (m->used * 2 > m->mask) {             /* grow */
        uint64_t ncap = (m->mask + 1) * 4, *old = m->set, ocap = m->mask + 1;
        m->set = (uint64_t *)malloc(ncap * sizeof(uint64_t));
        memset(m->set, 0xff, ncap * sizeof(uint64_t));
        m->mask = ncap - 1;
        m->used = 0;
        for (uint64_t i = 0; i < ocap; ++i) if (old[i] != ~0ULL) miner_put(m, old[i]);
        free(old);
    }
}

static void miner_search(miner_t *m, int r, int e, int goal, int *path, int depth, int rem_h, int rem_t)
{
    if (e == goal) {                                           /* rnnlogic.cpp:352-363 */
        if (depth == 1 && path[0] == r) return;                /* the trivial rule, erased at :532-539 */
        const int b = m->rel_bits, L = m->max_len;
        uint64_t key = ((uint64_t)r << (b * L + 3)) | ((uint64_t)depth << (b * L));
        for (int k = 0; k < depth; ++k) key |= (uint64_t)path[k] << (b * (L - 1 - k));
        miner_put(m, key);
        return;
    }
    if (depth == m->max_len) return;                           /* :364-367 */
    for (int k = m->adj_ptr[e]; k < m->adj_ptr[e + 1]; ++k) {  /* :370-381 (the set makes the visiting order irrelevant) */
        const int cr = m->adj_rel[k], cn = m->adj_dst[k];
        if (e == rem_h && cr == r && cn == rem_t) continue;
        path[depth] = cr;
        miner_search(m, r, cn, goal, path, depth + 1, rem_h, rem_t);
    }
}

static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

int64_t oracle_mine_rules(int64_t n_triples, const int32_t *tri, const int32_t *adj_ptr, const int32_t *adj_rel,
                          const int32_t *adj_dst, int max_len, int rel_bits, uint64_t *keys, int64_t cap)
{
    miner_t m = {adj_ptr, adj_rel, adj_dst, max_len, rel_bits, NULL, (1u << 16) - 1, 0};
    m.set = (uint64_t *)malloc((m.mask + 1) * sizeof(uint64_t));
    memset(m.set, 0xff, (m.mask + 1) * sizeof(uint64_t));
    int path[16];
    for (int64_t T = 0; T < n_triples; ++T)
        miner_search(&m, tri[3 * T + 1], tri[3 * T], tri[3 * T + 2], path, 0, tri[3 * T], tri[3 * T + 2]);
    int64_t n = 0;
    for (uint64_t i = 0; i <= m.mask; ++i)
        if (m.set[i] != ~0ULL) {
            if (n < cap) keys[n] = m.set[i];
            ++n;
        }
    qsort(keys, (size_t)(n < cap ? n : cap), sizeof(uint64_t), cmp_u64);
    free(m.set);
    return n;
}
