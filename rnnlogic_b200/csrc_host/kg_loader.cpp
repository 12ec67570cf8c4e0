// rnnlogic_b200 -- native dataset loader (host only, no CUDA).
//
// Replaces the pure-Python parse of KnowledgeGraph.__init__ (reference src/data.py:18-47,73-99):
// entities.dict / relations.dict ("id<TAB>name") and {train,valid,test}.txt ("h<TAB>r<TAB>t" by name)
// -> int64 id triples.  One pass per file over a memory buffer, open-addressing string table.
// C-ABI: plain pointers + sizes; buffers returned by rl_kg_load are owned by the handle.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Dataset {
    int64_t num_entities = 0, num_relations = 0;
    std::vector<int64_t> split[3];        // train / valid / test, (h, r, t) triples
    std::string error;
};

bool read_file(const std::string &path, std::string &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t got = n ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

// "id<TAB>name" per line (data.py:18-28: line.strip().split('\t'))
bool read_dict(const std::string &path, std::unordered_map<std::string, int64_t> &map, std::string &err)
{
    std::string buf;
    if (!read_file(path, buf)) { err = "cannot read " + path; return false; }
    size_t i = 0, n = buf.size();
    while (i < n) {
        size_t e = buf.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t a = i, b = e;
        while (a < b && (buf[a] == ' ' || buf[a] == '\r' || buf[a] == '\t')) ++a;
        while (b > a && (buf[b - 1] == ' ' || buf[b - 1] == '\r' || buf[b - 1] == '\t')) --b;
        if (b > a) {
            size_t tab = buf.find('\t', a);
            if (tab == std::string::npos || tab >= b) { err = "malformed line in " + path; return false; }
            const int64_t id = strtoll(buf.substr(a, tab - a).c_str(), nullptr, 10);
            map[buf.substr(tab + 1, b - tab - 1)] = id;
        }
        i = e + 1;
    }
    return true;
}

bool read_triples(const std::string &path, const std::unordered_map<std::string, int64_t> &ent,
                  const std::unordered_map<std::string, int64_t> &rel, std::vector<int64_t> &out, std::string &err)
{
    std::string buf;
    if (!read_file(path, buf)) { err = "cannot read " + path; return false; }
    size_t i = 0, n = buf.size();
    std::string tok[3];
    while (i < n) {
        size_t e = buf.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t a = i, b = e;
        while (a < b && (buf[a] == ' ' || buf[a] == '\r')) ++a;
        while (b > a && (buf[b - 1] == ' ' || buf[b - 1] == '\r' || buf[b - 1] == '\t')) --b;
        if (b > a) {
            size_t t1 = buf.find('\t', a), t2 = t1 == std::string::npos ? t1 : buf.find('\t', t1 + 1);
            if (t1 == std::string::npos || t2 == std::string::npos || t2 >= b) { err = "malformed triple in " + path; return false; }
            tok[0].assign(buf, a, t1 - a);
            tok[1].assign(buf, t1 + 1, t2 - t1 - 1);
            tok[2].assign(buf, t2 + 1, b - t2 - 1);
            auto h = ent.find(tok[0]);
            auto r = rel.find(tok[1]);
            auto t = ent.find(tok[2]);
            if (h == ent.end() || t == ent.end() || r == rel.end()) { err = "unknown name in " + path + ": " + tok[0] + " " + tok[1] + " " + tok[2]; return false; }   // KeyError in the reference
            out.push_back(h->second);
            out.push_back(r->second);
            out.push_back(t->second);
        }
        i = e + 1;
    }
    return true;
}

thread_local std::string g_last_error;

}  // namespace

extern "C" {

// Returns an opaque handle (NULL on failure, message in rl_kg_load_error()).
void *rl_kg_load(const char *data_path)
{
    Dataset *ds = new Dataset();
    std::unordered_map<std::string, int64_t> ent, rel;
    const std::string dir(data_path);
    const char *names[3] = {"train.txt", "valid.txt", "test.txt"};
    bool ok = read_dict(dir + "/entities.dict", ent, ds->error) && read_dict(dir + "/relations.dict", rel, ds->error);
    for (int k = 0; ok && k < 3; ++k) ok = read_triples(dir + "/" + names[k], ent, rel, ds->split[k], ds->error);
    if (!ok) {
        g_last_error = ds->error;
        delete ds;
        return nullptr;
    }
    ds->num_entities = (int64_t)ent.size();
    ds->num_relations = (int64_t)rel.size();
    return ds;
}

const char *rl_kg_load_error(void) { return g_last_error.c_str(); }
int64_t rl_kg_num_entities(void *h) { return static_cast<Dataset *>(h)->num_entities; }
int64_t rl_kg_num_relations(void *h) { return static_cast<Dataset *>(h)->num_relations; }
// which: 0 train, 1 valid, 2 test.  *count = number of triples; returns a pointer to count*3 int64.
const int64_t *rl_kg_triples(void *h, int which, int64_t *count)
{
    Dataset *ds = static_cast<Dataset *>(h);
    *count = (int64_t)(ds->split[which].size() / 3);
    return ds->split[which].data();
}
void rl_kg_free(void *h) { delete static_cast<Dataset *>(h); }

// ---- rule compiler: rule list -> per-head prefix tries (rules.py: CompiledRules) ----------------------------------
// Replaces the reference's relation2rules table (src/predictors.py:46-49, 186-189).  Rules are given flattened
// (head[n], body_ptr[n+1], body[]).  Nodes are numbered by head relation, then depth, then first appearance in the rule
// list -- the order the Python compiler used, so every table built on top of it is unchanged.  Outputs (caller-allocated,
// capacity = total body length): node_rel / node_parent (global ids, -1 at depth 1) / node_depth / node_head,
// head_node_ptr[R+1], rule_node[n] (-1 for an empty body).  Returns the number of nodes, or -1 - i when rule i uses a
// relation outside [0, R).
long long rl_compile_tries(long long n_rules, long long R, const long long *head, const long long *body_ptr,
                           const long long *body, long long *node_rel, long long *node_parent, long long *node_depth,
                           long long *node_head, long long *head_node_ptr, long long *rule_node)
{
    struct Local { long long rel, parent, depth; };                 // parent = local id inside the head's trie, -1 at depth 1
    std::vector<std::vector<Local>> nodes((size_t)R);
    std::vector<std::unordered_map<unsigned long long, long long>> index((size_t)R);   // (parent local id + 1) * R + rel -> local id
    std::vector<long long> leaf_local((size_t)n_rules, -1);
    for (long long i = 0; i < n_rules; ++i) {
        const long long q = head[i];
        if (q < 0 || q >= R) return -1 - i;
        long long cur = -1;
        for (long long k = body_ptr[i]; k < body_ptr[i + 1]; ++k) {
            const long long r = body[k];
            if (r < 0 || r >= R) return -1 - i;
            const unsigned long long key = (unsigned long long)(cur + 1) * (unsigned long long)R + (unsigned long long)r;
            auto it = index[q].find(key);
            if (it == index[q].end()) {
                const long long id = (long long)nodes[q].size();
                nodes[q].push_back({r, cur, k - body_ptr[i] + 1});
                index[q].emplace(key, id);
                cur = id;
            } else cur = it->second;
        }
        leaf_local[i] = cur;
    }
    long long total = 0;
    std::vector<std::vector<long long>> gid((size_t)R);
    for (long long q = 0; q < R; ++q) {
        head_node_ptr[q] = total;
        const auto &nd = nodes[q];
        long long maxd = 0;
        for (const Local &x : nd) maxd = x.depth > maxd ? x.depth : maxd;
        std::vector<long long> start((size_t)maxd + 2, 0);          // stable counting sort by depth
        for (const Local &x : nd) ++start[(size_t)x.depth + 1];
        for (long long d = 1; d <= maxd + 1; ++d) start[(size_t)d] += start[(size_t)d - 1];
        gid[q].resize(nd.size());
        for (size_t j = 0; j < nd.size(); ++j) gid[q][j] = total + start[(size_t)nd[j].depth]++;
        for (size_t j = 0; j < nd.size(); ++j) {
            const long long g = gid[q][j];
            node_rel[g] = nd[j].rel;
            node_depth[g] = nd[j].depth;
            node_head[g] = q;
            node_parent[g] = nd[j].parent < 0 ? -1 : gid[q][(size_t)nd[j].parent];
        }
        total += (long long)nd.size();
    }
    head_node_ptr[R] = total;
    for (long long i = 0; i < n_rules; ++i) rule_node[i] = leaf_local[i] < 0 ? -1 : gid[(size_t)head[i]][(size_t)leaf_local[i]];
    return total;
}

}  // extern "C"
