/*
 * sre_image.cpp -- breadth-first construction of the image automaton of a
 * lowered DFA (see sre_image.h).
 */
#include "sre_image.h"

#include <algorithm>
#include <unordered_map>

namespace {

typedef std::vector<uint16_t> set_t;

struct set_hash_t {
    size_t operator()(const set_t &v) const
    {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < v.size(); i++) {
            h = (h ^ v[i]) * 1099511628211ull;
        }
        return (size_t) h;
    }
};

}  // namespace

bool sre_build_image_automaton(const sre_dfa_t &dfa, uint32_t K, uint32_t max_narrow, uint32_t max_wide,
    sre_image_t &out)
{
    const uint32_t D = dfa.nstates, C = dfa.nclasses;
    if (D == 0 || C == 0 || K == 0 || K > 16 || D > 65534) {
        return false;
    }
    if (max_narrow + max_wide > 65000) {
        max_narrow = 65000 - max_wide;
    }
    std::unordered_map<set_t, uint32_t, set_hash_t> ids;
    std::vector<set_t> sets;
    uint32_t nnarrow = 0, nwide = 0;

    set_t top;
    for (uint32_t s = 0; s < D; s++) {
        if (s != dfa.acc) {
            top.push_back((uint16_t) s);
        }
    }
    /* U-state 0 is TOP even when the whole DFA is narrow (D - 1 <= K): "nothing
     * known" must stay distinguishable from a set that happens to hold every state */
    sets.push_back(top);
    nwide = 1;

    std::vector<uint8_t> mark(D, 0);
    set_t next;
    for (uint32_t u = 0; u < sets.size(); u++) {
        for (uint32_t c = 0; c < C; c++) {
            next.clear();
            const set_t &cur = sets[u];     /* (sets may grow below: re-read per class) */
            for (size_t i = 0; i < cur.size(); i++) {
                const uint16_t t = dfa.trans[(size_t) cur[i] * C + c];
                if (t != dfa.acc && !mark[t]) {
                    mark[t] = 1;
                    next.push_back(t);
                }
            }
            for (size_t i = 0; i < next.size(); i++) {
                mark[next[i]] = 0;
            }
            std::sort(next.begin(), next.end());
            uint32_t id;
            std::unordered_map<set_t, uint32_t, set_hash_t>::iterator it = ids.find(next);
            if (it != ids.end()) {
                id = it->second;
            } else {
                const bool narrow = next.size() <= K;
                if (narrow ? nnarrow >= max_narrow : nwide >= max_wide) {
                    id = 0;                 /* over budget: TOP, a superset */
                } else {
                    id = (uint32_t) sets.size();
                    ids[next] = id;
                    sets.push_back(next);
                    (narrow ? nnarrow : nwide)++;
                }
            }
            out.trans.resize((size_t) sets.size() * C, 0);
            out.trans[(size_t) u * C + c] = (uint16_t) id;
        }
    }

    out.nstates = (uint32_t) sets.size();
    out.nclasses = C;
    out.K = K;
    out.trans.resize((size_t) out.nstates * C, 0);
    out.cand.assign((size_t) out.nstates * K, 0xffff);
    out.ncand.assign(out.nstates, 0xff);
    for (uint32_t u = 1; u < out.nstates; u++) {
        if (sets[u].size() <= K) {
            out.ncand[u] = (uint8_t) sets[u].size();
            for (size_t i = 0; i < sets[u].size(); i++) {
                out.cand[(size_t) u * K + i] = sets[u][i];
            }
        }
    }
    return true;
}
