// rnnlogic_b200 -- rule discovery on the GPU.
// Reference: miner/rnnlogic.cpp:350-382 (KnowledgeGraph::rule_search) + :505-589 (RuleMiner::search_thread / search).
//
// For every train triple (h, r, t) the reference runs a depth-first search from h with the triple itself removed,
// records the relation sequence of every path that REACHES t within max_length hops (a path stops at its first visit
// of t; h == t yields the empty body), drops the trivial rule r <- r, and unions the bodies into a per-relation std::set.
//
// Here: one block per triple; the block's threads draw the first hops (the out-edges of h) from a shared work counter,
// so a hub head entity is spread over the whole block, and each thread walks the levels below its edge with an
// explicit stack.  A rule is one 64-bit key
//        head << (b*Lmax + 3) | length << (b*Lmax) | body[0] << (b*(Lmax-1)) | ... (left-aligned, b = bits of a relation id)
// whose ascending order IS the order of the reference's rule list: by head relation, then its std::set<Rule> order
// (length, body lexicographic; rnnlogic.cpp:118-133, 575-585).
// The union is a global open-addressing hash set of keys (atomicCAS); the host sorts the occupied slots.
#include "rl_device.cuh"

#define MINER_EMPTY 0xFFFFFFFFFFFFFFFFull
#define MINER_MAX_LEN 6

struct MinerArgs {
    int n_triples, max_len, rel_bits;
    const int32_t *tri;        // [n_triples][3] h, r, t
    const int32_t *adj_ptr;    // [N+1] out-edges of an entity (all relations)
    const int32_t *adj_rel;    // [E]
    const int32_t *adj_dst;    // [E]
    unsigned long long *table; // [cap] hash set of rule keys, cap a power of two
    unsigned long long cap_mask;
    int32_t *flags;            // [0] != 0: the table is full
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__device__ __forceinline__ void miner_insert(const MinerArgs &a, unsigned long long key)
{
    if (*reinterpret_cast<volatile int32_t *>(a.flags)) return;          // table already reported full: the host reruns with a larger one
    unsigned long long slot = mix64(key) & a.cap_mask;
    for (int probe = 0; probe < 256; ++probe) {
        const unsigned long long cur = a.table[slot];
        if (cur == key) return;
        if (cur == MINER_EMPTY) {
            const unsigned long long old = atomicCAS(a.table + slot, MINER_EMPTY, key);
            if (old == MINER_EMPTY || old == key) return;
        }
        slot = (slot + 1) & a.cap_mask;
    }
    a.flags[0] = 1;
}

__device__ __forceinline__ unsigned long long miner_key(const MinerArgs &a, int head, const int *path, int len)
{
    const int b = a.rel_bits, L = a.max_len;
    unsigned long long k = ((unsigned long long)head << (b * L + 3)) | ((unsigned long long)len << (b * L));
    for (int i = 0; i < len; ++i) k |= (unsigned long long)path[i] << (b * (L - 1 - i));
    return k;
}

// depth-first search below `e` (already `depth` hops from h, path[0..depth) filled, e != t, depth < max_len)
__device__ void miner_dfs(const MinerArgs &a, int h, int r, int t, int e, int depth, int *path)
{
    int node[MINER_MAX_LEN], pos[MINER_MAX_LEN];
    int sp = depth;
    node[sp] = e;
    pos[sp] = a.adj_ptr[e];
    while (sp >= depth) {
        const int cur = node[sp];
        if (pos[sp] >= a.adj_ptr[cur + 1]) { --sp; continue; }
        const int k = pos[sp]++;
        const int cr = a.adj_rel[k], cn = a.adj_dst[k];
        if (cur == h && cr == r && cn == t) continue;            // the triple itself is removed (rnnlogic.cpp:377)
        path[sp] = cr;
        if (cn == t) {                                           // reached the goal: a rule of sp + 1 hops; the search stops here
            if (!(sp == 0 && cr == r)) miner_insert(a, miner_key(a, r, path, sp + 1));   // r <- r is dropped (rnnlogic.cpp:532-539)
            continue;
        }
        if (sp + 1 < a.max_len) {
            ++sp;
            node[sp] = cn;
            pos[sp] = a.adj_ptr[cn];
        }
    }
}

__global__ void __launch_bounds__(256)
k_mine_rules(MinerArgs a)
{
    __shared__ int s_first, s_stop;
    for (int T = blockIdx.x; T < a.n_triples; T += gridDim.x) {
        if (threadIdx.x == 0) {
            s_first = 0;
            s_stop = *reinterpret_cast<volatile int32_t *>(a.flags);        // table reported full: the host reruns with a larger one
        }
        __syncthreads();
        if (s_stop) return;                                      // block-uniform
        const int h = a.tri[3 * T], r = a.tri[3 * T + 1], t = a.tri[3 * T + 2];
        if (h == t || a.max_len <= 0) {
            // e == goal at depth 0: the empty body (rnnlogic.cpp:352-363)
            if (h == t && threadIdx.x == 0) miner_insert(a, miner_key(a, r, nullptr, 0));
            __syncthreads();
            continue;
        }
        // level 1: the out-edges of h are drawn from a block-wide work counter; a thread goes on alone below its edge
        const int e0 = a.adj_ptr[h], e1 = a.adj_ptr[h + 1];
        int path[MINER_MAX_LEN];
        for (;;) {
            const int i = atomicAdd(&s_first, 1);
            if (e0 + i >= e1) break;
            const int cr = a.adj_rel[e0 + i], cn = a.adj_dst[e0 + i];
            if (cr == r && cn == t) continue;                    // the removed triple (cur == h here)
            path[0] = cr;
            if (cn == t) {
                if (cr != r) miner_insert(a, miner_key(a, r, path, 1));
                continue;
            }
            if (a.max_len > 1) miner_dfs(a, h, r, t, cn, 1, path);
        }
        __syncthreads();
    }
}

extern "C" {

/* Rule discovery (miner/rnnlogic.cpp:350-382, 505-589): every relation path of <= max_len hops from h to its first
 * visit of t, for every triple (h, r, t) of tri[n_triples][3] with that triple removed, without the trivial rule
 * r <- r.  adj_* = out-edges by source entity (DEVICE).  table[cap] (cap a power of two, filled with 0xFF bytes by the
 * caller) receives the distinct rule keys, see rl_miner.cu for the packing; flags[0] != 0 on return means the table
 * was too small.  rel_bits * (max_len + 1) + 3 must be <= 63. */
int rl_mine_rules(int32_t n_triples, const int32_t *tri, const int32_t *adj_ptr, const int32_t *adj_rel, const int32_t *adj_dst,
                  int32_t max_len, int32_t rel_bits, unsigned long long *table, int64_t cap, int32_t *flags, void *stream)
{
    if (!tri || !adj_ptr || !adj_rel || !adj_dst || !table || !flags) return rl_fail(RL_ERR_ARG, "rl_mine_rules: null argument");
    if (max_len < 0 || max_len > MINER_MAX_LEN || rel_bits <= 0 || rel_bits * (max_len + 1) + 3 > 63)
        return rl_fail(RL_ERR_ARG, "rl_mine_rules: max_len / rel_bits do not fit a 64-bit rule key");
    if (cap <= 0 || (cap & (cap - 1))) return rl_fail(RL_ERR_ARG, "rl_mine_rules: table capacity must be a power of two");
    if (n_triples <= 0) return RL_OK;
    MinerArgs a{n_triples, max_len, rel_bits, tri, adj_ptr, adj_rel, adj_dst, table, (unsigned long long)cap - 1ull, flags};
    const int grid = n_triples < 148 * 8 ? n_triples : 148 * 8;
    k_mine_rules<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    CHECK_LAUNCH("k_mine_rules");
    return RL_OK;
}

}  // extern "C"
