/*
 * sre_pdfa.cpp -- subset construction over ordered Pike thread lists (see
 * sre_pdfa.h).  The step below is the step of the closure-table Pike
 * (kernels/sre_pike_table.cu, host model oracle/lower_check.cpp:
 * lc_table_pike) without its next-byte pruning: a pruned thread is one that
 * dies on the next byte anyway, and an instruction that prunes one thread
 * prunes every thread parked on it, so the pruning changes no result.
 */
#include "sre_pdfa.h"

#include <map>
#include <string.h>

bool sre_build_pdfa(const sre_program_t *prog, const sre_closure_table_t &T, uint32_t max_states, sre_pdfa_t &out)
{
    const uint32_t np = T.npark;
    if (np == 0 || max_states > 0x7fff) {
        return false;
    }
    out.ctx_dep = T.ctx_dep;
    for (uint32_t P = 0; P < np; P++) {
        if (T.kind[P] >= 2) {
            return false;           /* look-ahead assertions: not a function of the byte alone */
        }
    }
    out.max_slots = T.max_slots;

    /* byte classes: bytes that the same parked instructions accept */
    {
        std::map<std::vector<uint8_t>, uint32_t> sigs;
        for (uint32_t b = 0; b < 256; b++) {
            std::vector<uint8_t> sig(T.nsets + 1);
            for (uint32_t k = 0; k < T.nsets; k++) {
                sig[k] = (T.accept[(size_t) k * 8 + (b >> 5)] >> (b & 31)) & 1;
            }
            /* the look-behind context a byte leaves: '\n' apart when closures depend on it */
            sig[T.nsets] = T.ctx_dep && b == '\n';
            std::map<std::vector<uint8_t>, uint32_t>::iterator it = sigs.find(sig);
            if (it == sigs.end()) {
                it = sigs.insert(std::make_pair(sig, (uint32_t) sigs.size())).first;
            }
            out.clsmap[b] = (uint8_t) it->second;
        }
        out.nclasses = (uint32_t) sigs.size();
    }
    const uint32_t C = out.nclasses;
    std::vector<uint32_t> rep(C, 0);
    for (int b = 255; b >= 0; b--) {
        rep[out.clsmap[b]] = (uint32_t) b;
    }
    auto accepts = [&](uint32_t P, uint32_t b) -> bool {
        return (T.accept[(size_t) T.acc_idx[P] * 8 + (b >> 5)] >> (b & 31)) & 1;
    };

    typedef std::vector<uint16_t> list_t;
    std::map<list_t, uint32_t> ids;
    std::vector<list_t> lists;
    lists.push_back(list_t());          /* state 0: the empty list */
    ids[list_t()] = 0;

    /* the start closures (closure of "pc 0", number np) by look-behind context */
    for (uint32_t v = 0; v < 3; v++) {
        if (v > 0 && !T.ctx_dep) {
            out.init[v] = out.init[0];
            out.init_mask_ofs[v] = out.init_mask_ofs[0];
            continue;
        }
        list_t init;
        std::vector<uint8_t> marks(np, 0);
        out.init_mask_ofs[v] = (uint32_t) out.init_mask.size();
        for (uint32_t e = T.ofs[v * (np + 2) + np]; e < T.ofs[v * (np + 2) + np + 1]; e++) {
            const uint32_t fp = T.ent[e];
            if (marks[fp]) {
                continue;
            }
            marks[fp] = 1;
            init.push_back((uint16_t) fp);
            out.init_mask.push_back(T.emask[e]);
        }
        if (init.empty() || init.size() > 255) {
            return false;
        }
        std::map<list_t, uint32_t>::iterator it = ids.find(init);
        if (it == ids.end()) {
            out.init[v] = (uint32_t) lists.size();
            ids[init] = out.init[v];
            lists.push_back(init);
        } else {
            out.init[v] = it->second;
        }
    }

    std::vector<uint8_t> marks(np);
    for (uint32_t s = 0; s < lists.size(); s++) {
        for (uint32_t c = 0; c < C; c++) {
            const list_t cur = lists[s];        /* copy: lists grows below */
            const uint32_t b = rep[c];
            /* closures appended after this byte see it as their look-behind context */
            const uint32_t vofs = (T.ctx_dep ? (b == '\n' ? 1u : 2u) : 0u) * (np + 2);
            list_t next;
            std::vector<uint8_t> parents;
            std::vector<uint32_t> masks;
            bool mev = false;
            uint8_t mpar = 0;
            uint32_t mmask = 0;
            uint16_t mreg = 0;
            memset(marks.data(), 0, np);
            for (size_t j = 0; j < cur.size() && !mev; j++) {
                const uint32_t P = cur[j];
                if (T.kind[P] == 1) {
                    /* a parked MATCH thread (only the start closure parks one): sre_vm_pike.c:535-553 */
                    mev = true;
                    mpar = (uint8_t) j;
                    mmask = 0;
                    mreg = T.regex[P];
                    break;
                }
                if (!accepts(P, b)) {
                    continue;
                }
                for (uint32_t e = T.ofs[vofs + P]; e < T.ofs[vofs + P + 1]; e++) {
                    const uint32_t fp = T.ent[e];
                    if (marks[fp]) {
                        continue;
                    }
                    marks[fp] = 1;
                    if (T.kind[fp] == 1) {
                        /* the closure reached MATCH: report and cut what has lower priority */
                        mev = true;
                        mpar = (uint8_t) j;
                        mmask = T.emask[e];
                        mreg = T.regex[fp];
                        break;
                    }
                    next.push_back((uint16_t) fp);
                    parents.push_back((uint8_t) j);
                    masks.push_back(T.emask[e]);
                }
            }
            if (next.size() > 255) {
                return false;
            }
            uint32_t id;
            std::map<list_t, uint32_t>::iterator it = ids.find(next);
            if (it == ids.end()) {
                id = (uint32_t) lists.size();
                if (id >= max_states) {
                    return false;
                }
                ids[next] = id;
                lists.push_back(next);
            } else {
                id = it->second;
            }
            out.trans.push_back((uint16_t) (id | (mev ? 0x8000u : 0u)));
            out.eofs.push_back((uint32_t) out.eparent.size());
            out.eparent.insert(out.eparent.end(), parents.begin(), parents.end());
            out.emask.insert(out.emask.end(), masks.begin(), masks.end());
            out.mparent.push_back(mpar);
            out.mmask.push_back(mmask);
            out.mregex.push_back(mreg);
            if (out.eparent.size() > (1u << 26)) {
                return false;
            }
        }
    }
    out.eofs.push_back((uint32_t) out.eparent.size());
    out.nstates = (uint32_t) lists.size();

    out.any_idx.assign(out.nstates, 0xff);
    out.eof_idx.assign(out.nstates, 0xff);
    out.eof_regex.assign(out.nstates, 0);
    out.list_ofs.assign(out.nstates + 1, 0);
    for (uint32_t s = 0; s < out.nstates; s++) {
        out.list_ofs[s] = (uint32_t) out.list_park.size();
        for (size_t j = 0; j < lists[s].size(); j++) {
            const uint32_t P = lists[s][j];
            out.list_park.push_back((uint16_t) P);
            if ((int32_t) P == T.p_any) {
                out.any_idx[s] = (uint8_t) j;
            }
            if (T.kind[P] == 1 && out.eof_idx[s] == 0xff) {
                out.eof_idx[s] = (uint8_t) j;
                out.eof_regex[s] = T.regex[P];
            }
        }
    }
    out.list_ofs[out.nstates] = (uint32_t) out.list_park.size();
    (void) prog;
    return true;
}
