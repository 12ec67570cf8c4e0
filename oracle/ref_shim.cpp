// TEST INFRASTRUCTURE ONLY.  C-ABI shim around the REFERENCE's own C++ path-count routine
// KnowledgeGraph::rule_destination (/root/reference/miner/rnnlogic.cpp:412-442), compiled
// from the reference sources where they lie (see oracle/Makefile) into oracle/_ref/.
// It is a second, independent pin for the oracle's grounding counts (SURVEY.md "Oracle B").
// No reference source is copied: this file only calls the class declared in rnnlogic.h.
#include "rnnlogic.h"

extern "C" {

void *ref_kg_new(const char *data_path)
{
    KnowledgeGraph *kg = new KnowledgeGraph();
    kg->read_data(const_cast<char *>(data_path));
    return kg;
}

void ref_kg_free(void *kg) { delete static_cast<KnowledgeGraph *>(kg); }

int ref_kg_entities(void *kg) { return static_cast<KnowledgeGraph *>(kg)->get_entity_size(); }

// Path counts of body[0..L) from entity e with the triple (rem_h, rem_r, rem_t) removed
// (pass rem_h = -1 for "no removal").  Writes up to cap (dest, count) pairs, returns how many
// destinations were reached.
int ref_rule_destination(void *kg, int e, int head, const int *body, int L,
                         int rem_h, int rem_r, int rem_t, int *dest, int *count, int cap)
{
    Rule rule;
    rule.r_head = head;
    for (int i = 0; i < L; ++i) rule.r_body.push_back(body[i]);
    Triplet removed;
    removed.h = rem_h; removed.r = rem_r; removed.t = rem_t;
    std::map<int, int> d2c;
    static_cast<KnowledgeGraph *>(kg)->rule_destination(e, rule, &d2c, removed);
    int n = 0;
    for (std::map<int, int>::iterator it = d2c.begin(); it != d2c.end(); ++it, ++n)
        if (n < cap) { dest[n] = it->first; count[n] = it->second; }
    return n;
}

}  // extern "C"

// ---- rule discovery: RuleMiner::search (/root/reference/miner/rnnlogic.cpp:505-589) ----
// Runs the reference's own miner on the loaded graph and writes the mined rules as consecutive records
// {head, length, body[0..length)} into out (at most cap ints).  Returns the number of ints of the full list.
extern "C" long long ref_mine_rules(void *kg, int max_length, int threads, int *out, long long cap)
{
    RuleMiner miner;
    miner.init_knowledge_graph(static_cast<KnowledgeGraph *>(kg));
    miner.search(max_length, 1.0, threads);
    std::vector<Rule> *rules = miner.get_logic_rules();
    const int R = static_cast<KnowledgeGraph *>(kg)->get_relation_size();
    long long n = 0;
    for (int r = 0; r < R; ++r)
        for (size_t i = 0; i < rules[r].size(); ++i) {
            const Rule &u = rules[r][i];
            if (n + 2 + (long long)u.r_body.size() <= cap) {
                out[n] = u.r_head;
                out[n + 1] = (int)u.r_body.size();
                for (size_t k = 0; k < u.r_body.size(); ++k) out[n + 2 + k] = u.r_body[k];
            }
            n += 2 + (long long)u.r_body.size();
        }
    return n;
}
