/*
 * sre_cuda_api.cu -- the C ABI of libsregex_cuda: the reference's executor
 * entry points (include/sregex/sregex.h) and the batch extension
 * (include/sregex_cuda.h), implemented on the CUDA kernels of ../kernels.
 *
 * No CPU fallback lives here: every *_exec entry point uploads its input (when
 * given a host buffer), launches kernels and downloads the verdict.  Without a
 * usable CUDA device the create/exec calls fail (NULL / SRE_ERROR) and say so
 * on stderr.
 */
#include <sregex_cuda.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../host/sre_internal.h"
#include "../kernels/sre_kernels.cuh"
#include "../lower/sre_lower.h"
#include "../lower/sre_closure.h"
#include "../lower/sre_image.h"
#include "../lower/sre_pdfa.h"
#include "../lower/sre_quirk.h"

namespace {

std::atomic<long>   g_launches(0);     /* statistics only */
thread_local char   g_err[256] = "";

const uint32_t MAX_DFA_STATES = 16384;    /* beyond: the NFA tier (the table is read from L2 when it exceeds shared memory) */
const uint32_t MAX_NFA_STATES = 4096;
const size_t   PIKE_SCRATCH_BUDGET = (size_t) 6 << 30;

int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "libsregex_cuda: %s\n", g_err);
    return SRE_ERROR;
}

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            return fail("%s failed: %s", #expr, cudaGetErrorString(e_));                 \
        }                                                                                \
    } while (0)

void count_launches(int n) { g_launches += n; }

/*
 * Per-call device scratch.  Nothing mutable lives in a program: every entry
 * point takes what it needs from the device's stream-ordered pool
 * (cudaMallocAsync) on the caller's stream and gives it back on the same stream
 * when the call returns, so one program can be used from any number of host
 * threads and CUDA streams at once.  The pool keeps freed blocks (release
 * threshold = never), so a steady caller pays for the allocation once.
 */
void pool_keep_memory()
{
    static std::once_flag once[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) {
        return;
    }
    std::call_once(once[dev], [dev]() {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    });
}

struct scratch_t {
    uint8_t      *p = nullptr;
    cudaStream_t  st = nullptr;
    scratch_t() {}
    scratch_t(const scratch_t &) = delete;
    scratch_t &operator=(const scratch_t &) = delete;
    ~scratch_t() { release(); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream)
    {
        release();
        pool_keep_memory();
        st = stream;
        void *q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, bytes ? bytes : 16, stream);
        p = static_cast<uint8_t *>(q);
        return e;
    }
    void release()
    {
        if (p) {
            cudaFreeAsync(p, st);
            p = nullptr;
        }
    }
};

bool device_ok()
{
    static int state = -1;
    if (state < 0) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        state = (e == cudaSuccess && n > 0) ? 1 : 0;
        if (!state) {
            cudaGetLastError();
        }
    }
    return state == 1;
}

/* host-side builder of one device blob holding every table of a program */
struct blob_t {
    std::vector<uint8_t> bytes;
    size_t add(const void *p, size_t n)
    {
        size_t ofs = (bytes.size() + 255) & ~(size_t) 255;
        bytes.resize(ofs + ((n + 15) & ~(size_t) 15), 0);
        if (n) {
            memcpy(&bytes[ofs], p, n);
        }
        return ofs;
    }
};

}  // namespace

/* Immutable after sre_cuda_program_create (apart from the two tuning words at
 * the end): all run-time state lives in per-call scratch or in contexts. */
struct sre_cuda_program_s {
    sre_program_t      *prog = nullptr;
    sre_lowered_t       low;
    sre_image_t         image;                  /* stream scan: image automaton of the DFA */
    uint8_t            *d_blob = nullptr;
    bool                has_dfa = false, has_nfa = false, has_image = false;
    sre_dev_dfa_t       dfa;
    sre_dev_image_t     img;
    sre_dev_nfa_t       nfa;
    sre_dev_nfa64_t     nfa64;                  /* nstates == 0: more than 64 lowered states */
    sre_dev_pike_t      pike;
    sre_dev_pdfa_t      pdfa;                   /* nstates == 0: no determinised Pike */
    uint32_t            nfa_shift = 0;
    /* byte values that leave the DFA start state (skip-scan tier), <= 4 kept */
    int                 nleave = 0;             /* -1: more than 4               */
    uint32_t            leave_pats[4] = { 0, 0, 0, 0 };
    /* tests: which Pike tier sre_cuda_pike_exec_lines may use / used last */
    std::atomic<int>    pike_tier_mode{0};
    std::atomic<int>    pike_last_tier{-1};
};

namespace {

void program_destroy(void *data)
{
    sre_cuda_program_t *cp = static_cast<sre_cuda_program_t *>(data);
    if (cp->prog) {
        cp->prog->lowered = nullptr;
    }
    cudaFree(cp->d_blob);
    delete cp;
}

int upload(sre_cuda_program_t *cp)
{
    const sre_program_t *prog = cp->prog;
    const sre_nfa_t &n = cp->low.nfa;
    blob_t b;
    size_t o_t256 = 0, o_tcls = 0, o_dcls = 0, o_fin = 0, o_h256 = 0, o_hcls = 0, o_hmap = 0;
    size_t o_itrans = 0, o_icand = 0, o_incand = 0, o_x256 = 0, o_x256m = 0;
    bool has_x256 = false, has_x256m = false;
    uint32_t xguess = 0;

    cp->has_dfa = cp->low.has_dfa;
    if (cp->has_dfa) {
        const sre_dfa_t &d = cp->low.dfa;
        if (!d.t256.empty()) {
            o_t256 = b.add(d.t256.data(), d.t256.size());
        }
        if (!d.h256.empty()) {
            o_h256 = b.add(d.h256.data(), d.h256.size());
        }
        if (!d.hcls.empty()) {
            o_hcls = b.add(d.hcls.data(), d.hcls.size() * 2);
            o_hmap = b.add(d.hclsmap, 256);
        }
        if (!d.t256.empty() && d.nstates <= 128) {
            /* text table (kernels/sre_text.cu): '\n' ends a line -- the verdict of the line (the
             * EOF step after its terminator) rides in bit 7 and the next line starts afresh */
            std::vector<uint8_t> x256((size_t) 256 * 256, 0);
            for (uint32_t st = 0; st < d.nstates; st++) {
                for (unsigned bv = 0; bv < 256; bv++) {
                    uint8_t e = d.t256[(size_t) st * 256 + bv];
                    if (bv == '\n') {
                        e = (uint8_t) (d.start | ((e == d.acc || d.fin[e]) ? 0x80u : 0u));
                    }
                    x256[(size_t) st * 256 + bv] = e;
                    x256[(size_t) (st + 128) * 256 + bv] = e;
                }
            }
            o_x256 = b.add(x256.data(), x256.size());
            has_x256 = true;
            if (d.nstates <= 64) {
                /* ... with the newline marked in bit 6 (rows r, r+64, r+128, r+192 alike) */
                std::vector<uint8_t> xm((size_t) 256 * 256, 0);
                for (uint32_t st = 0; st < d.nstates; st++) {
                    for (unsigned bv = 0; bv < 256; bv++) {
                        uint8_t e = x256[(size_t) st * 256 + bv];
                        if (bv == '\n') {
                            e |= 0x40;
                        }
                        for (uint32_t k = 0; k < 4; k++) {
                            xm[(size_t) (st + 64 * k) * 256 + bv] = e;
                        }
                    }
                }
                o_x256m = b.add(xm.data(), xm.size());
                has_x256m = true;
                /* the state the automaton idles in on ordinary text (k_text_verdicts enters a
                 * piece with it when a line is open there; a wrong guess only costs time): where
                 * most printable bytes, repeated, take the start state */
                uint32_t votes[64] = { 0 };
                for (unsigned bv = 0x20; bv < 0x7f; bv++) {
                    uint32_t st = d.start;
                    for (int k = 0; k < 64; k++) {
                        st = d.t256[(size_t) st * 256 + bv];
                    }
                    votes[st]++;
                }
                xguess = d.start;
                for (uint32_t st = 0; st < d.nstates; st++) {
                    if (votes[st] > votes[xguess]) {
                        xguess = st;
                    }
                }
            }
        }
        o_tcls = b.add(d.trans.data(), d.trans.size() * 2);
        o_dcls = b.add(d.clsmap, 256);
        o_fin = b.add(d.fin.data(), d.fin.size());
        /* image automaton for the stream scan: narrow sets up to 60000, 512 wide ones */
        cp->has_image = sre_build_image_automaton(d, SRE_STREAM_K, 60000, 512, cp->image);
        if (cp->has_image) {
            o_itrans = b.add(cp->image.trans.data(), cp->image.trans.size() * 2);
            o_icand = b.add(cp->image.cand.data(), cp->image.cand.size() * 2);
            o_incand = b.add(cp->image.ncand.data(), cp->image.ncand.size());
        }
    }

    /* NFA tables, every bitset row padded to WP = 32 * words-per-lane words */
    cp->has_nfa = n.nstates <= MAX_NFA_STATES;
    size_t o_ncls = 0, o_kind = 0, o_mv = 0, o_mt = 0, o_eof = 0, o_init = 0, o_shift = 0, o_follow = 0,
           o_rowidx = 0, o_any = 0, o_cmask = 0, o_mmask = 0;
    uint32_t WP = 32, nrows = 0;
    size_t o_mv64 = 0, o_mt64 = 0, o_fol64 = 0;
    uint64_t nfa64_any[3] = { 0, 0, 0 }, nfa64_shift = 0, nfa64_complex = 0, nfa64_init = 0, nfa64_eof = 0;
    bool has_nfa64 = false;
    if (cp->has_nfa) {
        const uint32_t wpl = (n.nwords + 31) / 32;
        WP = 32 * (wpl <= 1 ? 1 : wpl <= 2 ? 2 : 4);
        const uint32_t W = n.nwords;
        auto pad_rows = [&](const std::vector<uint32_t> &src, size_t rows) {
            std::vector<uint32_t> out(rows * WP, 0);
            for (size_t r = 0; r < rows; r++) {
                memcpy(&out[r * WP], &src[r * W], W * 4);
            }
            return out;
        };
        std::vector<uint32_t> mv = pad_rows(n.mv, n.nclasses), mt = pad_rows(n.mt, n.nclasses),
                              eofm = pad_rows(n.mt_eof, 1), init = pad_rows(n.init, 1),
                              shift = pad_rows(n.shift_mask, 1);
        std::vector<int32_t> rowidx((size_t) WP * 32, -1);
        std::vector<uint32_t> cmask(WP, 0), mmask(WP, 0), anyrow((size_t) n.nkinds * WP, 0);
        int32_t any_state = -1;
        for (uint32_t s = 0; s < n.nstates; s++) {
            if (n.state_pc[s] == 1 && n.state_allow[s] == 0x0f) {
                any_state = (int32_t) s;     /* the ".*?" prefix: always alive, consumes every byte */
            }
        }
        for (uint32_t s = 0; s < n.nstates; s++) {
            const bool is_shift = (n.shift_mask[s >> 5] >> (s & 31)) & 1;
            const bool is_match = prog->insts[n.state_pc[s]].opcode == SRE_OPCODE_MATCH;
            cp->nfa_shift += is_shift;
            if (is_match) {
                mmask[s >> 5] |= 1u << (s & 31);
            }
            if (!is_shift && !is_match && (int32_t) s != any_state) {
                rowidx[s] = (int32_t) nrows++;
                cmask[s >> 5] |= 1u << (s & 31);
            }
        }
        if (any_state < 0) {
            return fail("program without the .*? prefix state");
        }
        for (uint32_t k = 0; k < n.nkinds; k++) {
            memcpy(&anyrow[(size_t) k * WP], n.follow_row(k, (uint32_t) any_state), W * 4);
        }
        /* the any state must not also take the shift path */
        shift[any_state >> 5] &= ~(1u << (any_state & 31));
        std::vector<uint32_t> follow((size_t) n.nkinds * (nrows ? nrows : 1) * WP, 0);
        for (uint32_t k = 0; k < n.nkinds; k++) {
            for (uint32_t s = 0; s < n.nstates; s++) {
                if (rowidx[s] >= 0) {
                    memcpy(&follow[((size_t) k * nrows + rowidx[s]) * WP], n.follow_row(k, s), W * 4);
                }
            }
        }
        o_ncls = b.add(n.clsmap, 256);
        o_kind = b.add(n.cls_kind.data(), n.cls_kind.size());
        o_mv = b.add(mv.data(), mv.size() * 4);
        o_mt = b.add(mt.data(), mt.size() * 4);
        o_eof = b.add(eofm.data(), eofm.size() * 4);
        o_init = b.add(init.data(), init.size() * 4);
        o_shift = b.add(shift.data(), shift.size() * 4);
        o_follow = b.add(follow.data(), follow.size() * 4);
        o_rowidx = b.add(rowidx.data(), rowidx.size() * 4);
        o_any = b.add(anyrow.data(), anyrow.size() * 4);
        o_cmask = b.add(cmask.data(), cmask.size() * 4);
        o_mmask = b.add(mmask.data(), mmask.size() * 4);
        if (n.nstates <= 64) {
            /* 64-bit form for the thread-per-line kernel */
            auto w64 = [&](const uint32_t *row) -> uint64_t {
                return (uint64_t) row[0] | (W > 1 ? (uint64_t) row[1] << 32 : 0);
            };
            std::vector<uint64_t> mv64(n.nclasses), mt64(n.nclasses), fol64((size_t) n.nkinds * 64, 0);
            for (uint32_t c = 0; c < n.nclasses; c++) {
                mv64[c] = w64(&n.mv[(size_t) c * W]);
                mt64[c] = w64(&n.mt[(size_t) c * W]);
            }
            for (uint32_t k = 0; k < n.nkinds; k++) {
                for (uint32_t st = 0; st < n.nstates; st++) {
                    fol64[(size_t) k * 64 + st] = w64(n.follow_row(k, st));
                }
                nfa64_any[k] = w64(n.follow_row(k, (uint32_t) any_state));
            }
            nfa64_any[1] = n.nkinds > 1 ? nfa64_any[1] : nfa64_any[0];
            nfa64_any[2] = n.nkinds > 2 ? nfa64_any[2] : nfa64_any[0];
            nfa64_shift = w64(shift.data());
            nfa64_complex = w64(cmask.data());
            nfa64_init = w64(n.init.data());
            nfa64_eof = w64(n.mt_eof.data());
            o_mv64 = b.add(mv64.data(), mv64.size() * 8);
            o_mt64 = b.add(mt64.data(), mt64.size() * 8);
            o_fol64 = b.add(fol64.data(), fol64.size() * 8);
            has_nfa64 = true;
        }
    }

    /* Pike: the bytecode itself */
    std::vector<uint32_t> slot_ofs(prog->nregexes + 1, 0);
    for (sre_uint_t i = 0; i < prog->nregexes; i++) {
        slot_ofs[i + 1] = slot_ofs[i] + 2 * (uint32_t) (prog->multi_ncaps[i] + 1);
    }
    const size_t o_insts = b.add(prog->insts, (size_t) prog->len * sizeof(sre_instruction_t));
    const size_t o_ranges = b.add(prog->ranges, (size_t) prog->nranges * 2);
    const size_t o_leading = b.add(prog->leading, (size_t) prog->nleading * 4);
    const size_t o_slots = b.add(slot_ofs.data(), slot_ofs.size() * 4);
    /* the regex whose code region holds a pc: regions end at their MATCH */
    std::vector<uint16_t> pc_regex(prog->len + 1, 0);
    uint32_t max_slots = 2;
    {
        uint32_t r = 0;
        for (uint32_t pc = 0; pc < prog->len; pc++) {
            pc_regex[pc] = (uint16_t) (r < prog->nregexes ? r : prog->nregexes - 1);
            if (prog->insts[pc].opcode == SRE_OPCODE_MATCH) {
                r++;
            }
        }
        for (sre_uint_t i = 0; i < prog->nregexes; i++) {
            if (slot_ofs[i + 1] - slot_ofs[i] > max_slots) {
                max_slots = slot_ofs[i + 1] - slot_ofs[i];
            }
        }
    }
    const size_t o_pcre = b.add(pc_regex.data(), pc_regex.size() * 2);
    sre_closure_table_t clo;
    const bool has_clo = sre_build_closure_table(prog, 4096, clo);
    size_t o_cent = 0, o_cofs = 0, o_cacc = 0, o_ckind = 0, o_caidx = 0, o_creg = 0, o_cbent = 0, o_cbofs = 0;
    size_t o_cemask = 0, o_cbmask = 0;
    if (has_clo) {
        o_cent = b.add(clo.ent.data(), clo.ent.size() * 4);
        o_cemask = b.add(clo.emask.data(), clo.emask.size() * 4);
        o_cbmask = b.add(clo.bmask.data(), clo.bmask.size() * 4);
        o_cofs = b.add(clo.ofs.data(), clo.ofs.size() * 2);
        o_cacc = b.add(clo.accept.data(), clo.accept.size() * 4);
        o_ckind = b.add(clo.kind.data(), clo.kind.size());
        o_caidx = b.add(clo.acc_idx.data(), clo.acc_idx.size() * 2);
        o_creg = b.add(clo.regex.data(), clo.regex.size() * 2);
        o_cbent = b.add(clo.bent.data(), clo.bent.size() * 4);
        o_cbofs = b.add(clo.bofs.data(), clo.bofs.size() * 2);
    }
    /* the determinised Pike VM (up to 4096 thread lists of <= 255 threads) */
    sre_pdfa_t pd;
    const bool has_pd = has_clo && sre_build_pdfa(prog, clo, 4096, pd);
    size_t o_pcls = 0, o_ptrans = 0, o_peofs = 0, o_pent = 0, o_pmev = 0, o_peof = 0, o_pinit = 0;
    uint32_t pd_nent = 0, pd_init_any[4] = { 0xff, 0xff, 0xff, 0xff };
    size_t o_pent0 = 0, o_pmev0 = 0, o_peof0 = 0;
    if (has_pd) {
        /* device form of the provenance records (see sre_dev_pdfa_t) */
        const uint32_t C = pd.nclasses;
        std::vector<uint32_t> ent(2 * pd.eparent.size() + 2, 0), mev(2 * (size_t) pd.nstates * C, 0), eofv(pd.nstates);
        for (uint32_t st = 0; st < pd.nstates; st++) {
            for (uint32_t c = 0; c < C; c++) {
                const size_t t = (size_t) st * C + c;
                for (uint32_t i = pd.eofs[t]; i < pd.eofs[t + 1]; i++) {
                    ent[2 * (size_t) i] = pd.emask[i];
                    ent[2 * (size_t) i + 1] = pd.eparent[i] | (pd.eparent[i] == pd.any_idx[st] ? 0x100u : 0u);
                }
                mev[2 * t] = pd.mmask[t];
                mev[2 * t + 1] = pd.mparent[t] | (pd.mparent[t] == pd.any_idx[st] ? 0x100u : 0u)
                                 | ((uint32_t) pd.mregex[t] << 16);
            }
            eofv[st] = pd.eof_idx[st] | ((uint32_t) pd.eof_regex[st] << 16);
        }
        pd_nent = (uint32_t) pd.eparent.size();
        for (int v = 0; v < 4; v++) {
            pd_init_any[v] = pd.any_idx[pd.init[v]];
        }
        if (pd.lookahead) {
            std::vector<uint32_t> e0(pd.emask0);
            e0.push_back(0);
            o_pent0 = b.add(e0.data(), e0.size() * 4);
            o_pmev0 = b.add(pd.mmask0.data(), pd.mmask0.size() * 4);
            o_peof0 = b.add(pd.eof_mask0.data(), pd.eof_mask0.size() * 4);
        }
        o_pcls = b.add(pd.clsmap, 256);
        o_ptrans = b.add(pd.trans.data(), pd.trans.size() * 2);
        o_peofs = b.add(pd.eofs.data(), pd.eofs.size() * 4);
        o_pent = b.add(ent.data(), ent.size() * 4);
        o_pmev = b.add(mev.data(), mev.size() * 4);
        o_peof = b.add(eofv.data(), eofv.size() * 4);
        o_pinit = b.add(pd.init_mask.data(), pd.init_mask.size() * 4);
    }
    std::vector<uint32_t> start_ofs;
    static_assert(sizeof(sre_start_ent_t) == sizeof(sre_dev_start_t), "start entry layout");
    std::vector<sre_start_ent_t> start_ent;
    const bool has_start = sre_build_start_closure(prog, pc_regex, slot_ofs, start_ofs, start_ent);
    size_t o_sofs = 0, o_sent = 0;
    if (has_start) {
        o_sofs = b.add(start_ofs.data(), start_ofs.size() * 4);
        start_ent.push_back(sre_start_ent_t());       /* never an empty table */
        o_sent = b.add(start_ent.data(), start_ent.size() * sizeof(sre_start_ent_t));
    }

    /* the tables never straddle a 4 GiB boundary: kernels that chase a table through L1/L2
     * (k_dfa_lines_big) keep the high word of its addresses in a uniform register and do the
     * per-byte address arithmetic in 32 bits */
    {
        std::vector<void *> rejected;
        cudaError_t err = cudaSuccess;
        for (int attempt = 0; attempt < 8; attempt++) {
            void *ptr = nullptr;
            if ((err = cudaMalloc(&ptr, b.bytes.size() + 256)) != cudaSuccess) {
                break;
            }
            const uint64_t lo = reinterpret_cast<uint64_t>(ptr), hi = lo + b.bytes.size() + 255;
            if ((lo >> 32) == (hi >> 32)) {
                cp->d_blob = static_cast<uint8_t *>(ptr);
                break;
            }
            rejected.push_back(ptr);
        }
        for (void *ptr : rejected) {
            cudaFree(ptr);
        }
        if (err == cudaSuccess && cp->d_blob == nullptr) {
            return fail("uploading program tables failed: no allocation within one 4 GiB window");
        }
        if (err != cudaSuccess
            || cudaMemcpy(cp->d_blob, b.bytes.data(), b.bytes.size(), cudaMemcpyHostToDevice) != cudaSuccess)
        {
            return fail("uploading program tables failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    uint8_t *base = cp->d_blob;

    if (cp->has_dfa) {
        const sre_dfa_t &d = cp->low.dfa;
        cp->dfa.nstates = d.nstates;
        cp->dfa.nclasses = d.nclasses;
        cp->dfa.start = d.start;
        cp->dfa.acc = d.acc;
        cp->dfa.t256 = d.t256.empty() ? nullptr : base + o_t256;
        cp->dfa.tcls = reinterpret_cast<const uint16_t *>(base + o_tcls);
        cp->dfa.clsmap = base + o_dcls;
        cp->dfa.fin = base + o_fin;
        cp->dfa.h256 = d.h256.empty() ? nullptr : base + o_h256;
        cp->dfa.hcls = d.hcls.empty() ? nullptr : reinterpret_cast<const uint16_t *>(base + o_hcls);
        cp->dfa.hclsmap = d.hcls.empty() ? nullptr : base + o_hmap;
        cp->dfa.hncls = d.hncls;
        cp->dfa.x256 = has_x256 ? base + o_x256 : nullptr;
        cp->dfa.x256m = has_x256m ? base + o_x256m : nullptr;
        cp->dfa.xguess = xguess;
        if (cp->has_image) {
            cp->img.nstates = cp->image.nstates;
            cp->img.nclasses = cp->image.nclasses;
            cp->img.K = cp->image.K;
            cp->img.trans = reinterpret_cast<const uint16_t *>(base + o_itrans);
            cp->img.cand = reinterpret_cast<const uint16_t *>(base + o_icand);
            cp->img.ncand = base + o_incand;
        }
        cp->nleave = 0;
        if (!d.t256.empty()) {
            for (unsigned bv = 0; bv < 256 && cp->nleave >= 0; bv++) {
                if (d.t256[(size_t) d.start * 256 + bv] != d.start) {
                    if (cp->nleave == 4) {
                        cp->nleave = -1;
                    } else {
                        cp->leave_pats[cp->nleave++] = bv * 0x01010101u;
                    }
                }
            }
            for (int i = cp->nleave > 0 ? cp->nleave : 0; i < 4 && cp->nleave > 0; i++) {
                cp->leave_pats[i] = cp->leave_pats[0];
            }
        } else {
            cp->nleave = -1;
        }
    }
    if (cp->has_nfa) {
        cp->nfa.nstates = n.nstates;
        cp->nfa.nwords = WP;
        cp->nfa.nclasses = n.nclasses;
        cp->nfa.nkinds = n.nkinds;
        cp->nfa.clsmap = base + o_ncls;
        cp->nfa.cls_kind = base + o_kind;
        cp->nfa.mv = reinterpret_cast<const uint32_t *>(base + o_mv);
        cp->nfa.mt = reinterpret_cast<const uint32_t *>(base + o_mt);
        cp->nfa.mt_eof = reinterpret_cast<const uint32_t *>(base + o_eof);
        cp->nfa.init = reinterpret_cast<const uint32_t *>(base + o_init);
        cp->nfa.shift_mask = reinterpret_cast<const uint32_t *>(base + o_shift);
        cp->nfa.follow = reinterpret_cast<const uint32_t *>(base + o_follow);
        cp->nfa.rowidx = reinterpret_cast<const int32_t *>(base + o_rowidx);
        cp->nfa.nrows = nrows ? nrows : 1;
        cp->nfa.any_follow = reinterpret_cast<const uint32_t *>(base + o_any);
        cp->nfa.complex_mask = reinterpret_cast<const uint32_t *>(base + o_cmask);
        cp->nfa.match_mask = reinterpret_cast<const uint32_t *>(base + o_mmask);
        cp->nfa.match_lookahead = n.has_match_lookahead ? 1 : 0;
    }
    memset(&cp->nfa64, 0, sizeof(cp->nfa64));
    if (has_nfa64) {
        sre_dev_nfa64_t &q = cp->nfa64;
        q.nstates = n.nstates;
        q.nclasses = n.nclasses;
        q.nkinds = n.nkinds;
        q.clsmap = base + o_ncls;
        q.cls_kind = base + o_kind;
        q.mv = reinterpret_cast<const uint64_t *>(base + o_mv64);
        q.mt = reinterpret_cast<const uint64_t *>(base + o_mt64);
        q.follow = reinterpret_cast<const uint64_t *>(base + o_fol64);
        q.init = nfa64_init;
        q.mt_eof = nfa64_eof;
        q.shift_mask = nfa64_shift;
        q.complex_mask = nfa64_complex;
        q.any_follow[0] = nfa64_any[0];
        q.any_follow[1] = nfa64_any[1];
        q.any_follow[2] = nfa64_any[2];
        q.ncomplex = 0;
        for (uint32_t st = 0; st < 64; st++) {
            if ((nfa64_complex >> st) & 1) {
                if (q.ncomplex < 4) {
                    q.cidx[q.ncomplex] = (uint8_t) st;
                }
                q.ncomplex++;
            }
        }
        if (q.ncomplex > 4) {
            q.ncomplex = 0xffffffffu;
        }
    }

    memset(&cp->pdfa, 0, sizeof(cp->pdfa));
    if (has_pd) {
        sre_dev_pdfa_t &d = cp->pdfa;
        d.nstates = pd.nstates;
        d.nclasses = pd.nclasses;
        d.ctx_dep = 0;
        for (int v = 0; v < 4; v++) {
            d.init[v] = pd.init[v];
            d.init_any[v] = pd_init_any[v];
            d.init_mask_ofs[v] = pd.init_mask_ofs[v];
            d.ctx_dep |= pd.init[v] != pd.init[0] ? 1u : 0u;   /* the start list depends on the byte in front */
        }
        d.ent0 = pd.lookahead ? reinterpret_cast<const uint32_t *>(base + o_pent0) : nullptr;
        d.mev0 = pd.lookahead ? reinterpret_cast<const uint32_t *>(base + o_pmev0) : nullptr;
        d.eof0 = pd.lookahead ? reinterpret_cast<const uint32_t *>(base + o_peof0) : nullptr;
        d.max_slots = pd.max_slots;
        d.nent = pd_nent;
        d.clsmap = base + o_pcls;
        d.trans = reinterpret_cast<const uint16_t *>(base + o_ptrans);
        d.eofs = reinterpret_cast<const uint32_t *>(base + o_peofs);
        d.ent = reinterpret_cast<const uint2 *>(base + o_pent);
        d.mev = reinterpret_cast<const uint2 *>(base + o_pmev);
        d.eof = reinterpret_cast<const uint32_t *>(base + o_peof);
        d.init_mask = reinterpret_cast<const uint32_t *>(base + o_pinit);
    }

    static_assert(sizeof(sre_instruction_t) == sizeof(sre_dev_inst_t), "instruction layout");
    sre_dev_pike_t &pk = cp->pike;
    pk.len = prog->len;
    pk.nslots = slot_ofs[prog->nregexes];
    pk.nregexes = (uint32_t) prog->nregexes;
    pk.nleading = prog->nleading;
    pk.leading_byte = prog->leading_byte;
    pk.insts = reinterpret_cast<const sre_dev_inst_t *>(base + o_insts);
    pk.ranges = base + o_ranges;
    pk.leading = reinterpret_cast<const int32_t *>(base + o_leading);
    pk.slot_ofs = reinterpret_cast<const uint32_t *>(base + o_slots);
    pk.pc_regex = reinterpret_cast<const uint16_t *>(base + o_pcre);
    pk.clo_ent = has_clo ? reinterpret_cast<const uint32_t *>(base + o_cent) : nullptr;
    pk.clo_ofs = has_clo ? reinterpret_cast<const uint16_t *>(base + o_cofs) : nullptr;
    pk.clo_accept = has_clo ? reinterpret_cast<const uint32_t *>(base + o_cacc) : nullptr;
    pk.clo_kind = has_clo ? base + o_ckind : nullptr;
    pk.clo_nent = has_clo ? (uint32_t) clo.ent.size() : 0;
    pk.clo_npark = has_clo ? clo.npark : 0;
    pk.clo_emask = has_clo ? reinterpret_cast<const uint32_t *>(base + o_cemask) : nullptr;
    pk.clo_bmask = has_clo ? reinterpret_cast<const uint32_t *>(base + o_cbmask) : nullptr;
    pk.clo_accidx = has_clo ? reinterpret_cast<const uint16_t *>(base + o_caidx) : nullptr;
    pk.clo_regex = has_clo ? reinterpret_cast<const uint16_t *>(base + o_creg) : nullptr;
    pk.clo_bent = has_clo ? reinterpret_cast<const uint32_t *>(base + o_cbent) : nullptr;
    pk.clo_bofs = has_clo ? reinterpret_cast<const uint16_t *>(base + o_cbofs) : nullptr;
    pk.clo_nbent = has_clo ? (uint32_t) clo.bent.size() : 0;
    pk.clo_nsets = has_clo ? clo.nsets : 0;
    pk.clo_p_any = has_clo && clo.p_any >= 0 ? (uint32_t) clo.p_any : 0xffffffffu;
    pk.clo_has_hold = 0;
    for (uint8_t k : clo.kind) {
        if (k >= 2) {
            pk.clo_has_hold = 1;
        }
    }
    pk.clo_ctx_dep = clo.ctx_dep ? 1 : 0;
    pk.start_ofs = has_start ? reinterpret_cast<const uint32_t *>(base + o_sofs) : nullptr;
    pk.start_ent = has_start ? reinterpret_cast<const sre_dev_start_t *>(base + o_sent) : nullptr;
    pk.max_slots = max_slots;
    /* byte set of the leading instructions (sre_regex_compiler.c:123-241), for
     * the kernel's start-state shortcut */
    pk.quirk_possible = sre_quirk_bytes(prog, pk.quirk_single) ? 1u : 0u;
    memset(pk.leadset, 0, sizeof(pk.leadset));
    for (uint32_t i = 0; i < prog->nleading; i++) {
        const sre_instruction_t &in = prog->insts[prog->leading[i]];
        for (uint32_t b = 0; b < 256; b++) {
            bool hit = false;
            if (in.opcode == SRE_OPCODE_CHAR) {
                hit = (in.ch == b);
            } else {
                for (uint32_t j = 0; j < in.nranges; j++) {
                    const sre_vm_range_t &r = prog->ranges[in.v + j];
                    hit |= (b >= r.from && b <= r.to);
                }
                if (in.opcode == SRE_OPCODE_NOTIN) {
                    hit = !hit;
                }
            }
            if (hit) {
                pk.leadset[b >> 5] |= 1u << (b & 31);
            }
        }
    }
    pk.max_threads = 2 * prog->len + 16;
    pk.stack_cap = 2 * prog->len + 8;
    pk.ctx_stride = (sre_pike_ctx_bytes(pk.len, pk.nslots, pk.max_slots, pk.max_threads, pk.stack_cap) + 255)
                    & ~(uint64_t) 255;
    return SRE_OK;
}

cudaStream_t as_stream(void *s) { return static_cast<cudaStream_t>(s); }

/* pick the Thompson tier for a line batch; engine: SRE_CUDA_ENGINE_* | variant << 8 */
int thompson_dispatch(sre_cuda_program_t *cp, const uint8_t *dev_buf, const int64_t *dev_offsets,
    size_t nlines, size_t pitch, size_t linelen, int32_t *dev_rc, int engine, cudaStream_t st)
{
    int launches = 0;
    cudaError_t err;
    const bool aligned = dev_offsets == nullptr && (reinterpret_cast<uintptr_t>(dev_buf) & 15) == 0
                         && (pitch & 15) == 0 && linelen <= pitch && linelen < (1ull << 31);
    const int variant = (engine >> 8) & 0xff;
    engine &= 0xff;

    if (engine == SRE_CUDA_ENGINE_AUTO) {
        engine = !cp->has_dfa ? SRE_CUDA_ENGINE_NFA
               : !aligned ? SRE_CUDA_ENGINE_DFA_GENERIC
               : (cp->nleave >= 1 && cp->nleave <= 2 && linelen >= 128) ? SRE_CUDA_ENGINE_DFA_SKIP
               : SRE_CUDA_ENGINE_DFA_TILED;
    }
    switch (engine) {
    case SRE_CUDA_ENGINE_DFA_SKIP:
        if (!cp->has_dfa || !aligned || cp->nleave < 1) {
            return fail("DFA_SKIP engine needs a byte-table DFA whose start state is left by 1..4 byte "
                        "values, and 16-byte aligned fixed-pitch lines");
        }
        err = sre_launch_dfa_lines_skip(cp->dfa, dev_buf, nlines, pitch, linelen, dev_rc, cp->leave_pats,
                                        cp->nleave, variant, st, &launches);
        break;
    case SRE_CUDA_ENGINE_DFA_TILED:
        if (!cp->has_dfa || !aligned) {
            return fail("DFA_TILED engine needs a DFA and 16-byte aligned fixed-pitch lines");
        }
        err = sre_launch_dfa_lines(cp->dfa, dev_buf, nlines, pitch, linelen, dev_rc, variant, st, &launches);
        if (err == cudaErrorInvalidConfiguration) {
            /* table too large for shared memory next to the staging rings: tiled input, table
             * through L1/L2 */
            launches = 0;
            err = linelen >= 128 ? sre_launch_dfa_lines_big(cp->dfa, dev_buf, nlines, pitch, linelen, dev_rc, nullptr,
                                                            sre_gate_pack_t{ nullptr, nullptr, nullptr, 0 }, variant, st,
                                                            &launches)
                                 : sre_launch_dfa_ragged(cp->dfa, dev_buf, dev_offsets, nlines, pitch, linelen,
                                                         dev_rc, st, &launches);
        }
        break;
    case SRE_CUDA_ENGINE_DFA_GENERIC:
        if (!cp->has_dfa) {
            return fail("no DFA for this program (subset construction exceeded %u states)",
                        MAX_DFA_STATES);
        }
        err = sre_launch_dfa_ragged(cp->dfa, dev_buf, dev_offsets, nlines, pitch, linelen, dev_rc, st,
                                    &launches);
        break;
    case SRE_CUDA_ENGINE_NFA:
        if (!cp->has_nfa) {
            return fail("program has more than %u lowered states", MAX_NFA_STATES);
        }
        if (cp->nfa64.nstates != 0 && aligned && linelen >= 16) {
            /* up to 64 lowered states: one thread per line, the set in a 64-bit register */
            err = sre_launch_nfa64_lines(cp->nfa64, dev_buf, nlines, pitch, linelen, dev_rc, st, &launches);
            break;
        }
        /* fall through */
    case SRE_CUDA_ENGINE_NFA_WARP:
        if (!cp->has_nfa) {
            return fail("program has more than %u lowered states", MAX_NFA_STATES);
        }
        err = sre_launch_nfa_lines(cp->nfa, dev_buf, dev_offsets, nlines, pitch, linelen, nullptr, 1, 1,
                                   dev_rc, st, &launches);
        break;
    default:
        return fail("unknown engine %d", engine);
    }
    count_launches(launches);
    if (err != cudaSuccess) {
        return fail("Thompson kernel launch failed: %s", cudaGetErrorString(err));
    }
    return SRE_OK;
}

/* how many global-memory Pike contexts a call over nlines lines gets */
size_t pike_contexts(const sre_cuda_program_t *cp, size_t nlines)
{
    static long cap = -1;       /* SRE_PIKE_NCTX: concurrency override (tuning) */
    if (cap < 0) {
        const char *e = getenv("SRE_PIKE_NCTX");
        cap = e ? atol(e) : 0;
    }
    const size_t max_ctx = cap > 0 ? (size_t) cap : (size_t) 148 * 1024;
    size_t want = nlines < max_ctx ? nlines : max_ctx;
    const size_t by_budget = PIKE_SCRATCH_BUDGET / cp->pike.ctx_stride;
    if (want > by_budget) {
        want = by_budget;
    }
    return want ? want : 1;
}

/* contexts are interleaved in groups of 32 (sre_pike.cu: batch_base) */
size_t pike_scratch_bytes(const sre_cuda_program_t *cp, size_t nctx)
{
    return ((nctx + 31) / 32 * 32) * cp->pike.ctx_stride;
}

}  // namespace

/* one reduced stream part: its records (all levels) stay on the device until
 * the part has been resolved with its true entry state */
struct sre_cuda_stream_scan_s {
    sre_cuda_program_t *cp = nullptr;
    const uint8_t      *buf = nullptr;
    size_t              len = 0;
    cudaStream_t        st = nullptr;
    scratch_t           mem;
    sre_stream_ws_t     ws;
    int                 top = 0;
    /* device results: [0] fixed piece, [1] exit state | entry of the ACC piece, [2] match offset */
    unsigned long long *d_res = nullptr;
};

namespace {

/* size the levels of the stream scan and carve the workspace */
int stream_scan_alloc(sre_cuda_stream_scan_t *sc)
{
    const size_t piece = sre_stream_piece_bytes(), fan = sre_stream_fan();
    const size_t fs = SRE_STREAM_FN_BYTES;
    sre_stream_ws_t &ws = sc->ws;
    size_t cnt = sc->len / piece + ((sc->len % piece) || sc->len == 0 ? 1 : 0), total = 0;
    int l = 0;
    for (;; l++) {
        ws.count[l] = cnt;
        total += (cnt * fs + 255) & ~(size_t) 255;
        if (cnt <= fan || l == 3) {
            break;
        }
        cnt = (cnt + fan - 1) / fan;
    }
    if (ws.count[l] > fan) {
        return fail("stream too long for the 4-level scan");
    }
    sc->top = l;
    for (int k = l + 1; k < 4; k++) {
        ws.count[k] = 0;
        ws.fn[k] = nullptr;
    }
    total += 256;
    CUDA_TRY(sc->mem.alloc(total, sc->st));
    uint8_t *p = sc->mem.p;
    for (int k = 0; k <= l; k++) {
        ws.fn[k] = p;
        p += (ws.count[k] * fs + 255) & ~(size_t) 255;
    }
    ws.first_acc = reinterpret_cast<unsigned long long *>(p);
    sc->d_res = reinterpret_cast<unsigned long long *>(p + 64);
    return SRE_OK;
}

/* the top level's records composed into the part's one record (host) */
int stream_root_record(sre_cuda_stream_scan_t *sc, uint8_t *host_fn)
{
    std::vector<uint8_t> top(sc->ws.count[sc->top] * SRE_STREAM_FN_BYTES);
    CUDA_TRY(cudaMemcpyAsync(top.data(), sc->ws.fn[sc->top], top.size(), cudaMemcpyDeviceToHost, sc->st));
    CUDA_TRY(cudaStreamSynchronize(sc->st));
    sre_stream_fn_identity(host_fn);
    for (size_t j = 0; j < sc->ws.count[sc->top]; j++) {
        sre_stream_fn_compose(host_fn, &top[j * SRE_STREAM_FN_BYTES]);
    }
    return SRE_OK;
}

int stream_reduce(sre_cuda_program_t *cp, const uint8_t *dev_buf, size_t len, const uint8_t *dev_halo,
    uint32_t entry, cudaStream_t st, sre_cuda_stream_scan_t **out)
{
    *out = nullptr;
    if (!cp->has_dfa || !cp->has_image) {
        return fail("the stream scan needs the determinised program (this one exceeded %u DFA states)",
                    MAX_DFA_STATES);
    }
    if (reinterpret_cast<uintptr_t>(dev_buf) & 15) {
        return fail("stream buffer must be 16-byte aligned");
    }
    if (entry != SRE_STREAM_UNKNOWN && entry >= cp->dfa.nstates) {
        return fail("bad stream state");
    }
    sre_cuda_stream_scan_t *sc = new (std::nothrow) sre_cuda_stream_scan_t();
    if (sc == nullptr) {
        return fail("out of memory");
    }
    sc->cp = cp;
    sc->buf = dev_buf;
    sc->len = len;
    sc->st = st;
    if (stream_scan_alloc(sc) != SRE_OK) {
        delete sc;
        return SRE_ERROR;
    }
    int launches = 0;
    cudaError_t err = sre_launch_dfa_stream_reduce(cp->dfa, cp->img, dev_buf, len, dev_halo, entry, sc->ws, st,
                                                   &launches);
    count_launches(launches);
    if (err != cudaSuccess) {
        delete sc;
        return fail("stream kernels failed: %s", cudaGetErrorString(err));
    }
    *out = sc;
    return SRE_OK;
}

/*
 * With the true entry state: repair the unresolved pieces (normally none), walk
 * down to the first match.  One device round trip when nothing needs repair.
 */
int stream_resolve(sre_cuda_stream_scan_t *sc, uint32_t entry, uint32_t *exit_state, int64_t *match_offset,
    size_t *repaired)
{
    sre_cuda_program_t *cp = sc->cp;
    struct { unsigned long long fixed; uint32_t out[2]; long long off; } h;
    size_t rounds = 0;
    for (int burst = 1;; burst = 16) {
        int launches = 0;
        cudaError_t err = cudaSuccess;
        for (int i = 0; i < burst && err == cudaSuccess; i++) {
            err = sre_launch_dfa_stream_fix(cp->dfa, sc->buf, sc->len, entry, sc->ws, sc->d_res, sc->st, &launches);
        }
        if (err == cudaSuccess) {
            err = sre_launch_dfa_stream_walk(cp->dfa, cp->img, sc->buf, sc->len, entry, sc->ws,
                                             reinterpret_cast<uint32_t *>(sc->d_res + 1),
                                             reinterpret_cast<long long *>(sc->d_res + 2), sc->st, &launches);
        }
        count_launches(launches);
        if (err != cudaSuccess) {
            return fail("stream kernels failed: %s", cudaGetErrorString(err));
        }
        CUDA_TRY(cudaMemcpyAsync(&h, sc->d_res, sizeof(h), cudaMemcpyDeviceToHost, sc->st));
        CUDA_TRY(cudaStreamSynchronize(sc->st));
        if (h.fixed == ~0ull) {
            break;              /* the last repair round found nothing left: the walk is valid */
        }
        rounds += burst;
        if (rounds > sc->ws.count[0] + 16) {
            return fail("stream scan: repair does not terminate");
        }
    }
    if (h.out[0] == 0xffffffffu) {
        return fail("stream scan: a piece does not know the state it was entered in");
    }
    *exit_state = h.out[0];
    *match_offset = entry == cp->dfa.acc ? 0 : h.off;
    if (repaired) {
        *repaired = rounds;
    }
    return SRE_OK;
}

}  // namespace

/* ======================================================================== *
 * batch extension (include/sregex_cuda.h)
 * ======================================================================== */

extern "C" {

SRE_API int sre_cuda_device_available(void) { return device_ok() ? 1 : 0; }
SRE_API const char *sre_cuda_last_error(void) { return g_err; }

SRE_API int
sre_cuda_index_lines(const uint8_t *dev_buf, size_t len, int64_t *dev_offsets, size_t max_lines, size_t *nlines,
    void *stream)
{
    if ((dev_buf == NULL && len != 0) || dev_offsets == NULL || nlines == NULL) {
        return fail("NULL buffer, offsets or nlines");
    }
    cudaStream_t st = as_stream(stream);
    const size_t need = sre_lines_workspace_bytes(len);
    scratch_t ws;
    CUDA_TRY(ws.alloc(need, st));
    unsigned long long *w = reinterpret_cast<unsigned long long *>(ws.p);
    int launches = 0;
    cudaError_t err = sre_launch_index_lines(dev_buf, len, dev_offsets, max_lines, w, st, &launches);
    count_launches(launches);
    unsigned long long found = 0;
    if (err == cudaSuccess) {
        err = cudaMemcpyAsync(&found, w + need / sizeof(unsigned long long) - 1, sizeof(found),
                              cudaMemcpyDeviceToHost, st);
    }
    if (err == cudaSuccess) {
        err = cudaStreamSynchronize(st);
    }
    if (err != cudaSuccess) {
        return fail("line index failed: %s", cudaGetErrorString(err));
    }
    *nlines = (size_t) found;
    return SRE_OK;
}

/* grep: see include/sregex_cuda.h */
SRE_API int
sre_cuda_thompson_exec_text(sre_cuda_program_t *cp, const uint8_t *dev_buf, size_t len, int64_t *dev_offsets,
    int32_t *dev_rc, size_t max_lines, size_t *nlines, void *stream)
{
    if (cp == NULL || nlines == NULL || (dev_buf == NULL && len != 0) || (dev_rc == NULL && max_lines != 0)) {
        return fail("NULL program, buffer, rc or nlines");
    }
    cudaStream_t st = as_stream(stream);
    int launches = 0;
    unsigned long long found = 0;
    if (cp->has_dfa && cp->dfa.x256 != nullptr && (reinterpret_cast<uintptr_t>(dev_buf) & 15) == 0) {
        /* one pass: pieces through the TMA pipeline, line structure folded into the table */
        scratch_t ws;
        CUDA_TRY(ws.alloc(sre_text_workspace_bytes(len), st));
        cudaError_t err = sre_launch_text(cp->dfa, dev_buf, len, dev_rc, dev_offsets, max_lines, ws.p, st, &launches);
        count_launches(launches);
        if (err == cudaSuccess) {
            err = cudaMemcpyAsync(&found, ws.p + sre_text_count_offset(len), sizeof(found), cudaMemcpyDeviceToHost,
                                  st);
        }
        if (err == cudaSuccess) {
            err = cudaStreamSynchronize(st);
        }
        if (err != cudaSuccess) {
            return fail("text kernels failed: %s", cudaGetErrorString(err));
        }
        *nlines = (size_t) found;
        return SRE_OK;
    }
    /* larger automata / unaligned buffers: line index, then the ragged tier */
    scratch_t own;
    int64_t *offs = dev_offsets;
    if (offs == NULL) {
        CUDA_TRY(own.alloc((max_lines + 1) * sizeof(int64_t), st));
        offs = reinterpret_cast<int64_t *>(own.p);
    }
    size_t n = 0;
    if (sre_cuda_index_lines(dev_buf, len, offs, max_lines, &n, stream) != SRE_OK) {
        return SRE_ERROR;
    }
    *nlines = n;
    const size_t run = n < max_lines ? n : max_lines;
    return run ? sre_cuda_thompson_exec_ragged(cp, dev_buf, offs, run, dev_rc, SRE_CUDA_ENGINE_AUTO, stream) : SRE_OK;
}

SRE_API void
sre_cuda_program_set_pike_tier(sre_cuda_program_t *cp, int mode)
{
    if (cp) {
        cp->pike_tier_mode = mode;
    }
}

SRE_API int
sre_cuda_program_last_pike_tier(sre_cuda_program_t *cp)
{
    return cp ? cp->pike_last_tier.load() : -1;
}

SRE_API long sre_cuda_launch_count(int reset)
{
    return reset ? g_launches.exchange(0) : g_launches.load();
}

SRE_API sre_cuda_program_t *
sre_cuda_program_create(sre_program_t *prog)
{
    if (prog == NULL || prog->magic != SRE_PROGRAM_MAGIC) {
        fail("not a program compiled by this library");
        return NULL;
    }
    /* one lowering per program, also when several threads ask at once */
    static std::mutex create_lock;
    std::lock_guard<std::mutex> guard(create_lock);
    if (prog->lowered) {
        return static_cast<sre_cuda_program_t *>(prog->lowered);
    }
    if (!device_ok()) {
        fail("no usable CUDA device (libsregex_cuda has no CPU fallback)");
        return NULL;
    }
    sre_cuda_program_t *cp = new (std::nothrow) sre_cuda_program_t();
    if (cp == NULL) {
        return NULL;
    }
    cp->prog = prog;
    if (sre_lower_program(prog, MAX_DFA_STATES, &cp->low) != SRE_OK || upload(cp) != SRE_OK) {
        cp->prog = nullptr;
        program_destroy(cp);
        return NULL;
    }
    if (sre_pool_add_cleanup(prog->pool, program_destroy, cp) != SRE_OK) {
        cp->prog = nullptr;
        program_destroy(cp);
        return NULL;
    }
    prog->lowered = cp;
    return cp;
}

SRE_API int
sre_cuda_program_info(sre_cuda_program_t *cp, sre_cuda_info_t *info)
{
    if (cp == NULL || info == NULL) {
        return SRE_ERROR;
    }
    memset(info, 0, sizeof(*info));
    info->prog_len = cp->prog->len;
    info->nfa_states = cp->low.nfa.nstates;
    info->nfa_classes = cp->low.nfa.nclasses;
    info->nfa_kinds = cp->low.nfa.nkinds;
    info->nfa_shift_states = cp->nfa_shift;
    info->dfa_states = cp->has_dfa ? cp->low.dfa.nstates : 0;
    info->dfa_classes = cp->has_dfa ? cp->low.dfa.nclasses : 0;
    info->dfa_byte_table = cp->has_dfa && cp->dfa.t256 != nullptr;
    info->dfa_leave_bytes = cp->has_dfa && cp->nleave > 0 ? (uint32_t) cp->nleave : 0;
    info->nregexes = (uint32_t) cp->prog->nregexes;
    info->pike_slots = cp->pike.nslots;
    info->pike_ctx_bytes = cp->pike.ctx_stride;
    info->dfa_start = cp->has_dfa ? cp->dfa.start : 0;
    info->dfa_acc = cp->has_dfa ? cp->dfa.acc : 0;
    info->image_states = cp->has_image ? cp->image.nstates : 0;
    return SRE_OK;
}

SRE_API int
sre_cuda_thompson_exec_lines(sre_cuda_program_t *cp, const uint8_t *dev_buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *dev_rc, int engine, void *stream)
{
    if (cp == NULL) {
        return fail("NULL program");
    }
    if (nlines == 0) {
        return SRE_OK;
    }
    return thompson_dispatch(cp, dev_buf, nullptr, nlines, pitch, linelen, dev_rc, engine,
                             as_stream(stream));
}

SRE_API int
sre_cuda_thompson_exec_ragged(sre_cuda_program_t *cp, const uint8_t *dev_buf,
    const int64_t *dev_offsets, size_t nlines, int32_t *dev_rc, int engine, void *stream)
{
    if (cp == NULL || dev_offsets == NULL) {
        return fail("NULL program or offsets");
    }
    if (nlines == 0) {
        return SRE_OK;
    }
    if ((engine & 0xff) == SRE_CUDA_ENGINE_DFA_TILED) {
        engine = SRE_CUDA_ENGINE_DFA_GENERIC;
    }
    return thompson_dispatch(cp, dev_buf, dev_offsets, nlines, 0, 0, dev_rc, engine, as_stream(stream));
}

SRE_API int
sre_cuda_pike_exec_lines(sre_cuda_program_t *cp, const uint8_t *dev_buf, const int64_t *dev_offsets,
    size_t nlines, size_t pitch, size_t linelen, const int32_t *dev_select, int32_t *dev_rc,
    int64_t *dev_ovec, size_t ovec_slots, void *stream)
{
    if (cp == NULL) {
        return fail("NULL program");
    }
    if (nlines == 0) {
        return SRE_OK;
    }
    int launches = 0;
    cudaStream_t st = as_stream(stream);
    const int32_t *start = nullptr;

    /*
     * No caller-supplied gate: run the determinised program first.  It yields
     * the Thompson verdict (lines that cannot match are skipped) and, per
     * line, the offset after which no earlier-started thread is alive, where
     * the Pike search may begin.
     */
    /* per-line workspace: gate | start hint | packed line list | its count */
    const size_t half = (nlines * 4 + 255) & ~(size_t) 255;
    scratch_t line_ws;
    CUDA_TRY(line_ws.alloc(3 * half + 256, st));
    const bool aligned = dev_offsets == nullptr && (reinterpret_cast<uintptr_t>(dev_buf) & 15) == 0
                         && (pitch & 15) == 0 && linelen <= pitch && linelen < (1ull << 31);
    const bool tiled_hint = cp->has_dfa && cp->dfa.h256 != nullptr && aligned;
    /* packed list of the lines to run + its count (behind the gate and the hints) */
    if (nlines > 0xffffffffull) {
        return fail("too many lines for one call");
    }
    uint32_t *list = reinterpret_cast<uint32_t *>(line_ws.p + 2 * half);
    uint32_t *count = reinterpret_cast<uint32_t *>(line_ws.p + 3 * half);
    bool packed = false;        /* the gate kernel packed the list itself */
    if (tiled_hint || (cp->has_dfa && cp->dfa.hcls != nullptr)) {
        int32_t *gate = reinterpret_cast<int32_t *>(line_ws.p);
        int32_t *hint = reinterpret_cast<int32_t *>(line_ws.p + half);
        /* our own gate on tiled lines: the verdicts go straight to dev_rc (final for the lines that
         * do not match) and the kernel appends the matching lines to the list as it goes */
        const bool tiled = tiled_hint || (aligned && linelen >= 128);
        packed = tiled && dev_select == nullptr;
        sre_gate_pack_t pack = { nullptr, nullptr, nullptr, 0 };
        if (packed) {
            CUDA_TRY(cudaMemsetAsync(count, 0, 4, st));
            gate = dev_rc;
            pack = sre_gate_pack_t{ list, count, dev_ovec, (uint32_t) ovec_slots };
        }
        cudaError_t e = tiled_hint
            ? sre_launch_dfa_lines_hint(cp->dfa, dev_buf, nlines, pitch, linelen, gate, hint,
                                        (cp->nleave >= 1 && cp->nleave <= 2 && linelen >= 128) ? cp->leave_pats
                                                                                                : nullptr,
                                        cp->nleave, pack, st, &launches)
            : (aligned && linelen >= 128)
            ? sre_launch_dfa_lines_big(cp->dfa, dev_buf, nlines, pitch, linelen, gate, hint, pack, 0, st, &launches)
            : sre_launch_dfa_generic_hint(cp->dfa, dev_buf, dev_offsets, nlines, pitch, linelen, gate, hint, st,
                                          &launches);
        if (e != cudaSuccess) {
            count_launches(launches);
            return fail("hint kernel launch failed: %s", cudaGetErrorString(e));
        }
        /* a caller-supplied gate is honoured as given; the start hints are
         * valid for every line either way */
        if (dev_select == nullptr) {
            dev_select = gate;
        }
        start = hint;
    }

    cudaError_t err;
    /* pack the selected lines; the rows of the others are final right away
     * (rc from the gate, ovector all -1 = all 0xff bytes) */
    sre_line_list_t lines = { nullptr, nullptr };
    if (dev_select != nullptr) {
        if (!packed) {
            /* (a gate kernel that packs the list sets the rows of the other lines itself) */
            if (dev_ovec != nullptr && ovec_slots != 0) {
                CUDA_TRY(cudaMemsetAsync(dev_ovec, 0xff, nlines * ovec_slots * sizeof(int64_t), st));
            }
            CUDA_TRY(cudaMemsetAsync(count, 0, 4, st));
            err = sre_launch_pike_compact(dev_select, nlines, dev_rc, list, count, st, &launches);
            if (err != cudaSuccess) {
                count_launches(launches);
                return fail("compaction kernel launch failed: %s", cudaGetErrorString(err));
            }
        }
        lines.list = list;
        lines.count = count;
    }
    /* the table kernel's bookkeeping lives next to the packed-list count */
    sre_pike_work_t *work = reinterpret_cast<sre_pike_work_t *>(line_ws.p + 3 * half + 64);
    /* closure-table kernel: list capacities of its two passes.  Small lists
     * first (more resident warps), then the lines that needed more; a set of
     * regexes can have as many live threads as members share a prefix. */
    static int ek1 = -1, eh1 = 2, ek2 = 0, eh2 = 0;
    if (ek1 < 0) {              /* SRE_PIKE_TABLE_K="K1,H1[,K2,H2]" (tuning) */
        const char *e = getenv("SRE_PIKE_TABLE_K");
        int k1e = 0;
        if (e) {
            sscanf(e, "%d,%d,%d,%d", &k1e, &eh1, &ek2, &eh2);
        }
        ek1 = k1e;
    }
    const bool big = cp->pike.nregexes > 1 || cp->pike.clo_npark > 64;
    const int k1 = ek1 > 0 ? ek1 : (big ? 12 : 4);
    int k2 = ek2 > 0 ? ek2 : (big ? 32 : 16);
    /* no look-ahead assertion in the program: no pending closures to hold */
    const int h1 = cp->pike.clo_has_hold ? (ek1 > 0 ? eh1 : 2) : 0;
    const int h2 = cp->pike.clo_has_hold ? (ek2 > 0 ? eh2 : 4) : 0;
    if (cp->pike.clo_npark && (uint32_t) k2 > cp->pike.clo_npark) {
        k2 = (int) cp->pike.clo_npark;
    }
    /* wide capture vectors: shrink the retry lists until a block fits in shared memory */
    while (k2 > k1 && !sre_pike_table_applicable(cp->pike, dev_offsets, linelen, k2, h2)) {
        k2 = k2 - 4 > k1 ? k2 - 4 : k1;
    }
    const int tier_mode = cp->pike_tier_mode.load();
    const bool table_ok = sre_pike_table_applicable(cp->pike, dev_offsets, linelen, k2 > k1 ? k2 : k1, h2)
                          && linelen < (1ull << 31);
    /* the determinised Pike VM first when the program has one; the closure-table
     * kernel (with its larger lists) re-runs the lines whose match outlives the ring */
    /* (SRE_PDFA_LOOKAHEAD=0 keeps programs with look-ahead assertions on the closure-table kernel) */
    static const int la_ok = [] {
        const char *e = getenv("SRE_PDFA_LOOKAHEAD");
        return e ? atoi(e) : 1;
    }();
    const bool use_lineage = tier_mode == 0 && sre_pike_lineage_applicable(cp->pdfa, linelen) && table_ok
                             && (cp->pdfa.ent0 == nullptr || la_ok);
    const bool use_table = !use_lineage && table_ok && (tier_mode == 0 || tier_mode == 3);
    const bool use_small = !use_lineage && !use_table && sre_pike_small_applicable(cp->pike)
                           && linelen < (1ull << 31) && tier_mode != 1;
    /* global-memory contexts of k_pike_lines: one per concurrent line when it
     * does all the work, a few thousand when it only re-runs what a
     * shared-memory tier gave up on */
    const size_t nctx = pike_contexts(cp, (use_lineage || use_table || use_small) ? (nlines < 16384 ? nlines : 16384)
                                                                                   : nlines);
    scratch_t pike_scratch;
    CUDA_TRY(pike_scratch.alloc(pike_scratch_bytes(cp, nctx), st));
    if (use_lineage) {
        cp->pike_last_tier = 3;
        /* rows of a gated batch were set to -1 by the memset above */
        err = sre_launch_pike_lineage(cp->pdfa, dev_buf, dev_offsets, nlines, pitch, linelen, lines, start,
                                      dev_rc, dev_ovec, (uint32_t) ovec_slots, lines.list != nullptr && !packed, work,
                                      st, &launches);
        if (err == cudaSuccess) {
            err = sre_launch_pike_table(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines, start,
                                        dev_rc, dev_ovec, (uint32_t) ovec_slots, k2 > k1 ? k2 : k1, h2, 1, work, st,
                                        &launches);
        }
        if (err == cudaSuccess) {
            err = sre_launch_pike_lines(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines,
                                        start, dev_rc, dev_ovec, (uint32_t) ovec_slots, pike_scratch.p,
                                        nctx, 1, st, &launches, &work->given_up[1]);
        }
    } else
    if (use_table) {
        /* the general kernel re-runs what the table kernel gave up on */
        cp->pike_last_tier = 0;
        err = sre_launch_pike_table(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines, start,
                                    dev_rc, dev_ovec, (uint32_t) ovec_slots, k1, h1, 0, work, st, &launches);
        const bool two = (k1 < k2 || h1 < h2);
        if (err == cudaSuccess && two) {
            err = sre_launch_pike_table(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines, start,
                                        dev_rc, dev_ovec, (uint32_t) ovec_slots, k2, h2, 1, work, st, &launches);
        }
        if (err == cudaSuccess) {
            err = sre_launch_pike_lines(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines,
                                        start, dev_rc, dev_ovec, (uint32_t) ovec_slots, pike_scratch.p,
                                        nctx, 1, st, &launches, &work->given_up[two ? 1 : 0]);
        }
    } else if (use_small) {
        /* shared-memory kernel first; the general kernel re-runs what it gave up on */
        cp->pike_last_tier = 2;
        err = sre_launch_pike_small(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines, start,
                                    dev_rc, dev_ovec, (uint32_t) ovec_slots, st, &launches);
        if (err == cudaSuccess) {
            err = sre_launch_pike_lines(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines,
                                        start, dev_rc, dev_ovec, (uint32_t) ovec_slots, pike_scratch.p,
                                        nctx, 1, st, &launches);
        }
    } else {
        cp->pike_last_tier = 1;
        err = sre_launch_pike_lines(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines, start,
                                    dev_rc, dev_ovec, (uint32_t) ovec_slots, pike_scratch.p, nctx, 0, st,
                                    &launches);
    }
    /* the reference's first-byte prefilter can misfire on some (program, line) pairs and report a
     * later match than the leftmost one (lower/sre_quirk.h): those lines are
     * replayed to the letter, from offset 0, by the general kernel */
    if (err == cudaSuccess && cp->pike.quirk_possible) {
        uint32_t *qcount = reinterpret_cast<uint32_t *>(line_ws.p + 3 * half + 128);
        err = sre_launch_pike_quirk_mark(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines, dev_rc,
                                         dev_ovec, (uint32_t) ovec_slots, qcount, st, &launches);
        if (err == cudaSuccess) {
            err = sre_launch_pike_lines(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen, lines, start,
                                        dev_rc, dev_ovec, (uint32_t) ovec_slots, pike_scratch.p, nctx, 2, st,
                                        &launches, qcount);
        }
    }
    count_launches(launches);
    if (err != cudaSuccess) {
        return fail("Pike kernel launch failed: %s", cudaGetErrorString(err));
    }
    return SRE_OK;
}

SRE_API int
sre_cuda_pike_exec_lines_all(sre_cuda_program_t *cp, const uint8_t *dev_buf, const int64_t *dev_offsets,
    size_t nlines, size_t pitch, size_t linelen, size_t max_matches, int32_t *dev_count, int64_t *dev_spans,
    int32_t *dev_ids, void *stream)
{
    if (cp == NULL || max_matches == 0) {
        return fail("NULL program or max_matches == 0");
    }
    if (nlines == 0) {
        return SRE_OK;
    }
    cudaStream_t st = as_stream(stream);
    const size_t nctx = pike_contexts(cp, nlines);
    scratch_t pike_scratch;
    CUDA_TRY(pike_scratch.alloc(pike_scratch_bytes(cp, nctx), st));
    int launches = 0;
    cudaError_t err = sre_launch_pike_lines_all(cp->pike, dev_buf, dev_offsets, nlines, pitch, linelen,
                                                (uint32_t) max_matches, dev_count, dev_spans, dev_ids,
                                                pike_scratch.p, nctx < nlines ? nctx : nlines, st, &launches);
    count_launches(launches);
    if (err != cudaSuccess) {
        return fail("Pike kernel launch failed: %s", cudaGetErrorString(err));
    }
    return SRE_OK;
}

SRE_API sre_cuda_stream_scan_t *
sre_cuda_thompson_stream_reduce(sre_cuda_program_t *cp, const uint8_t *dev_buf, size_t len,
    const uint8_t *dev_halo, uint32_t entry_state, uint8_t *host_fn, void *stream)
{
    if (cp == NULL) {
        fail("NULL program");
        return NULL;
    }
    const uint32_t entry = entry_state == SRE_CUDA_STATE_INIT ? cp->dfa.start
                         : entry_state == SRE_CUDA_STATE_UNKNOWN ? SRE_STREAM_UNKNOWN : entry_state;
    sre_cuda_stream_scan_t *sc = nullptr;
    if (stream_reduce(cp, dev_buf, len, dev_halo, entry, as_stream(stream), &sc) != SRE_OK) {
        return NULL;
    }
    if (host_fn != NULL && stream_root_record(sc, host_fn) != SRE_OK) {
        delete sc;
        return NULL;
    }
    return sc;
}

SRE_API int
sre_cuda_thompson_stream_resolve(sre_cuda_stream_scan_t *scan, uint32_t entry_state, uint32_t *exit_state,
    int64_t *first_match_offset, uint8_t *host_fn)
{
    if (scan == NULL || exit_state == NULL) {
        return fail("NULL scan");
    }
    const uint32_t entry = entry_state == SRE_CUDA_STATE_INIT ? scan->cp->dfa.start : entry_state;
    if (entry >= scan->cp->dfa.nstates) {
        return fail("bad stream state");
    }
    int64_t off = -1;
    if (stream_resolve(scan, entry, exit_state, &off, nullptr) != SRE_OK) {
        return SRE_ERROR;
    }
    if (first_match_offset) {
        *first_match_offset = off;
    }
    if (host_fn != NULL) {
        return stream_root_record(scan, host_fn);
    }
    return SRE_OK;
}

SRE_API void
sre_cuda_thompson_stream_free(sre_cuda_stream_scan_t *scan)
{
    delete scan;
}

SRE_API uint32_t
sre_cuda_stream_fn_apply(const uint8_t *fn, uint32_t state)
{
    const uint32_t r = sre_stream_fn_apply(fn, state);
    return r == SRE_STREAM_UNKNOWN ? SRE_CUDA_STATE_UNKNOWN : r;
}

SRE_API int
sre_cuda_thompson_exec_stream(sre_cuda_program_t *cp, const uint8_t *dev_buf, size_t len,
    size_t chunk_bytes, unsigned eof, uint32_t *state_io, int64_t *match_chunk, void *stream)
{
    if (cp == NULL || state_io == NULL) {
        return fail("NULL program or state");
    }
    if (!cp->has_dfa) {
        return fail("the stream scan needs the determinised program (this one exceeded %u DFA states)",
                    MAX_DFA_STATES);
    }
    const uint32_t entry = *state_io == SRE_CUDA_STATE_INIT ? cp->dfa.start : *state_io;
    sre_cuda_stream_scan_t *sc = nullptr;
    if (stream_reduce(cp, dev_buf, len, nullptr, entry, as_stream(stream), &sc) != SRE_OK) {
        return SRE_ERROR;
    }
    uint32_t exit_state = 0;
    int64_t off = -1;
    const int r = stream_resolve(sc, entry, &exit_state, &off, nullptr);
    delete sc;
    if (r != SRE_OK) {
        return SRE_ERROR;
    }
    *state_io = exit_state;

    const size_t nchunks = chunk_bytes ? (len + chunk_bytes - 1) / chunk_bytes : 1;
    if (exit_state == cp->dfa.acc) {
        if (match_chunk) {
            *match_chunk = (entry == cp->dfa.acc || off < 0 || !chunk_bytes)
                               ? 0 : (int64_t) ((size_t) off / chunk_bytes);
        }
        return SRE_OK;
    }
    if (eof) {
        if (cp->low.dfa.fin[exit_state]) {
            if (match_chunk) {
                *match_chunk = nchunks ? (int64_t) nchunks - 1 : 0;
            }
            return SRE_OK;
        }
        return SRE_DECLINED;
    }
    return SRE_AGAIN;
}

/*
 * The same over a stream that lives in HOST memory: slices of slice_bytes are
 * copied to the device on one CUDA stream while the previous slice is scanned
 * on another; the automaton state is carried from slice to slice exactly as
 * *state_io carries it from call to call.  Stops at the slice in which the
 * match is seen, like a caller of the reference would stop feeding chunks.
 */
SRE_API int
sre_cuda_thompson_exec_stream_host(sre_cuda_program_t *cp, const uint8_t *host_buf, size_t len,
    size_t chunk_bytes, unsigned eof, uint32_t *state_io, int64_t *match_chunk, size_t slice_bytes)
{
    if (cp == NULL || state_io == NULL || (host_buf == NULL && len != 0)) {
        return fail("NULL program, state or buffer");
    }
    if (slice_bytes == 0) {
        slice_bytes = (size_t) 256 << 20;
    }
    if (chunk_bytes != 0) {
        /* slices hold whole chunks, so that a chunk index is slice base + local index */
        slice_bytes = slice_bytes < chunk_bytes ? chunk_bytes : slice_bytes / chunk_bytes * chunk_bytes;
    }
    slice_bytes = (slice_bytes + 15) & ~(size_t) 15;
    if (chunk_bytes % 16 != 0 && len > slice_bytes) {
        return fail("chunk_bytes must be a multiple of 16 for multi-slice host streams");
    }
    cudaStream_t copy_st = nullptr, scan_st = nullptr;
    cudaEvent_t copied[2] = { nullptr, nullptr };
    scratch_t slice[2];
    int rc = SRE_ERROR;
    cudaError_t err = cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&scan_st, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && err == cudaSuccess; i++) {
        err = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming);
        if (err == cudaSuccess) {
            err = slice[i].alloc(slice_bytes + 16, scan_st);
        }
    }
    if (err == cudaSuccess) {
        /* the slices are used by both streams: make the allocations visible to the copy stream */
        err = cudaStreamSynchronize(scan_st);
    }
    const size_t nslices = len == 0 ? 1 : (len + slice_bytes - 1) / slice_bytes;
    auto enqueue_copy = [&](size_t k) -> cudaError_t {
        const size_t at = k * slice_bytes, n = len - at < slice_bytes ? len - at : slice_bytes;
        cudaError_t e = n ? cudaMemcpyAsync(slice[k & 1].p, host_buf + at, n, cudaMemcpyHostToDevice, copy_st)
                          : cudaSuccess;
        return e == cudaSuccess ? cudaEventRecord(copied[k & 1], copy_st) : e;
    };
    if (err == cudaSuccess) {
        err = enqueue_copy(0);
    }
    for (size_t k = 0; k < nslices && err == cudaSuccess; k++) {
        const size_t at = k * slice_bytes, n = len - at < slice_bytes ? len - at : slice_bytes;
        const bool last = k + 1 == nslices;
        /* slice k+1 travels while slice k is scanned (its buffer was released when
         * the scan of slice k-1 returned) */
        if (!last && (err = enqueue_copy(k + 1)) != cudaSuccess) {
            break;
        }
        if ((err = cudaStreamWaitEvent(scan_st, copied[k & 1], 0)) != cudaSuccess) {
            break;
        }
        int64_t mc = -1;
        rc = sre_cuda_thompson_exec_stream(cp, slice[k & 1].p, n, chunk_bytes, last ? eof : 0, state_io, &mc,
                                           scan_st);
        if (rc == SRE_OK && match_chunk) {
            *match_chunk = (chunk_bytes ? (int64_t) (at / chunk_bytes) : 0) + (mc > 0 ? mc : 0);
        }
        if (rc != SRE_AGAIN) {
            break;
        }
    }
    if (err != cudaSuccess) {
        fail("host stream scan failed: %s", cudaGetErrorString(err));
        rc = SRE_ERROR;
    }
    /* copies still in flight (early stop) must be over before the slices go back to the pool */
    if (copy_st) cudaStreamSynchronize(copy_st);
    if (scan_st) cudaStreamSynchronize(scan_st);
    slice[0].release();
    slice[1].release();
    if (scan_st) cudaStreamSynchronize(scan_st);
    for (int i = 0; i < 2; i++) {
        if (copied[i]) cudaEventDestroy(copied[i]);
    }
    if (copy_st) cudaStreamDestroy(copy_st);
    if (scan_st) cudaStreamDestroy(scan_st);
    return rc;
}

/* ---- batched streaming Pike contexts ------------------------------------------ */

}  /* extern "C" */

struct sre_cuda_pike_streams_s {
    sre_cuda_program_t *cp = nullptr;
    size_t              nstreams = 0;
    uint8_t            *d_ctx = nullptr;
};

extern "C" {

SRE_API sre_cuda_pike_streams_t *
sre_cuda_pike_streams_create(sre_cuda_program_t *cp, size_t nstreams, void *stream)
{
    if (cp == NULL || nstreams == 0) {
        fail("NULL program or no streams");
        return NULL;
    }
    sre_cuda_pike_streams_t *h = new (std::nothrow) sre_cuda_pike_streams_t();
    if (h == NULL) {
        return NULL;
    }
    h->cp = cp;
    h->nstreams = nstreams;
    int launches = 0;
    if (cudaMalloc(&h->d_ctx, nstreams * cp->pike.ctx_stride) != cudaSuccess
        || sre_launch_pike_streams_init(cp->pike, h->d_ctx, nstreams, as_stream(stream), &launches) != cudaSuccess)
    {
        fail("creating %zu Pike contexts failed: %s", nstreams, cudaGetErrorString(cudaGetLastError()));
        cudaFree(h->d_ctx);
        delete h;
        return NULL;
    }
    count_launches(launches);
    return h;
}

SRE_API int
sre_cuda_pike_streams_exec(sre_cuda_pike_streams_t *h, const uint8_t *dev_buf, const int64_t *dev_offsets,
    const uint8_t *dev_eof, unsigned eof_all, int64_t *dev_out, size_t ovec_slots, void *stream)
{
    if (h == NULL || dev_offsets == NULL || dev_out == NULL) {
        return fail("NULL handle, offsets or output");
    }
    int launches = 0;
    cudaError_t err = sre_launch_pike_streams(h->cp->pike, h->d_ctx, h->nstreams, dev_buf, dev_offsets, dev_eof,
                                              eof_all != 0, dev_out, (uint32_t) ovec_slots, as_stream(stream),
                                              &launches);
    count_launches(launches);
    return err == cudaSuccess ? SRE_OK : fail("Pike stream kernel launch failed: %s", cudaGetErrorString(err));
}

SRE_API void
sre_cuda_pike_streams_free(sre_cuda_pike_streams_t *h)
{
    if (h) {
        cudaFree(h->d_ctx);
        delete h;
    }
}

SRE_API int
sre_cuda_dfa_fin(sre_cuda_program_t *cp, uint32_t state)
{
    if (cp == NULL || !cp->has_dfa || state >= cp->low.dfa.nstates) {
        return 0;
    }
    return cp->low.dfa.fin[state];
}

/* ---- host-buffer conveniences --------------------------------------------- */

SRE_API int
sre_cuda_thompson_exec_lines_host(sre_cuda_program_t *cp, const uint8_t *host_buf, size_t nlines,
    size_t pitch, size_t linelen, int32_t *host_rc, int engine)
{
    if (cp == NULL) {
        return fail("NULL program");
    }
    if (nlines == 0) {
        return SRE_OK;
    }
    cudaStream_t st = cudaStreamPerThread;
    const size_t in_bytes = (nlines - 1) * pitch + linelen, in_pad = (in_bytes + 255) & ~(size_t) 255;
    scratch_t io;
    CUDA_TRY(io.alloc(in_pad + nlines * 4 + 256, st));
    int32_t *d_rc = reinterpret_cast<int32_t *>(io.p + in_pad);
    CUDA_TRY(cudaMemcpyAsync(io.p, host_buf, in_bytes, cudaMemcpyHostToDevice, st));
    if (thompson_dispatch(cp, io.p, nullptr, nlines, pitch, linelen, d_rc, engine, st) != SRE_OK) {
        return SRE_ERROR;
    }
    CUDA_TRY(cudaMemcpyAsync(host_rc, d_rc, nlines * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SRE_OK;
}

SRE_API int
sre_cuda_pike_exec_lines_host(sre_cuda_program_t *cp, const uint8_t *host_buf, size_t nlines,
    size_t pitch, size_t linelen, int gate_with_thompson, int32_t *host_rc, int64_t *host_ovec,
    size_t ovec_slots)
{
    if (cp == NULL) {
        return fail("NULL program");
    }
    if (nlines == 0) {
        return SRE_OK;
    }
    cudaStream_t st = cudaStreamPerThread;
    const size_t in_bytes = (nlines - 1) * pitch + linelen, in_pad = (in_bytes + 255) & ~(size_t) 255;
    const size_t rc_pad = (nlines * 4 + 255) & ~(size_t) 255;
    scratch_t io;
    CUDA_TRY(io.alloc(in_pad + 2 * rc_pad + nlines * ovec_slots * 8 + 256, st));
    int32_t *d_sel = reinterpret_cast<int32_t *>(io.p + in_pad);
    int32_t *d_rc = reinterpret_cast<int32_t *>(io.p + in_pad + rc_pad);
    int64_t *d_ov = reinterpret_cast<int64_t *>(io.p + in_pad + 2 * rc_pad);
    CUDA_TRY(cudaMemcpyAsync(io.p, host_buf, in_bytes, cudaMemcpyHostToDevice, st));
    if (gate_with_thompson
        && thompson_dispatch(cp, io.p, nullptr, nlines, pitch, linelen, d_sel, SRE_CUDA_ENGINE_AUTO, st) != SRE_OK)
    {
        return SRE_ERROR;
    }
    if (sre_cuda_pike_exec_lines(cp, io.p, nullptr, nlines, pitch, linelen,
                                 gate_with_thompson ? d_sel : nullptr, d_rc, d_ov, ovec_slots, st) != SRE_OK)
    {
        return SRE_ERROR;
    }
    CUDA_TRY(cudaMemcpyAsync(host_rc, d_rc, nlines * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(host_ovec, d_ov, nlines * ovec_slots * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SRE_OK;
}

}  /* extern "C" */

/* ======================================================================== *
 * the reference's executor API (include/sregex/sregex.h)
 * ======================================================================== */

struct sre_vm_thompson_ctx_s {
    sre_cuda_program_t  *cp;
    uint8_t             *d_mem;         /* state (WP words) | rc | input       */
    size_t               d_cap;         /* input capacity                      */
    uint32_t             state_words;
    bool                 started;
};

struct sre_vm_pike_ctx_s {
    sre_cuda_program_t  *cp;
    sre_int_t           *ovector;
    size_t               ovec_slots;
    uint8_t             *d_ctx;         /* persistent Pike context             */
    int64_t             *d_out;
    uint8_t             *d_in;
    size_t               d_in_cap;
    std::vector<int64_t> *h_out;
    sre_int_t            pending[2];
    bool                 started;       /* an exec call has been made */
};

struct sre_vm_thompson_code_s {
    sre_cuda_program_t  *cp;
};

namespace {

void thompson_ctx_cleanup(void *data)
{
    sre_vm_thompson_ctx_t *ctx = static_cast<sre_vm_thompson_ctx_t *>(data);
    cudaFree(ctx->d_mem);
    ctx->d_mem = nullptr;
}

void pike_ctx_cleanup(void *data)
{
    sre_vm_pike_ctx_t *ctx = static_cast<sre_vm_pike_ctx_t *>(data);
    cudaFree(ctx->d_ctx);
    cudaFree(ctx->d_out);
    cudaFree(ctx->d_in);
    delete ctx->h_out;
    ctx->d_ctx = nullptr;
    ctx->d_out = nullptr;
    ctx->d_in = nullptr;
    ctx->h_out = nullptr;
}

const size_t THOMPSON_HDR = 4096 * 4 + 256;     /* state words + rc */

/* The classic entry points have no stream argument: each host thread works on
 * its own (per-thread default) stream, so contexts driven from different
 * threads do not serialise on the legacy default stream. */
const cudaStream_t CLASSIC_STREAM = cudaStreamPerThread;

int thompson_reserve(sre_vm_thompson_ctx_t *ctx, size_t len)
{
    if (ctx->d_mem && ctx->d_cap >= len) {
        return SRE_OK;
    }
    uint8_t *fresh = nullptr;
    const size_t cap = len < 4096 ? 4096 : (len + 255) & ~(size_t) 255;
    CUDA_TRY(cudaMalloc(&fresh, THOMPSON_HDR + cap));
    if (ctx->d_mem) {
        cudaMemcpy(fresh, ctx->d_mem, THOMPSON_HDR, cudaMemcpyDeviceToDevice);
        cudaFree(ctx->d_mem);
    }
    ctx->d_mem = fresh;
    ctx->d_cap = cap;
    return SRE_OK;
}

}  // namespace

extern "C" {

/* reference: sre_vm_thompson_create_ctx, sre_vm_thompson.c:25-60 */
SRE_API sre_vm_thompson_ctx_t *
sre_vm_thompson_create_ctx(sre_pool_t *pool, sre_program_t *prog)
{
    sre_cuda_program_t *cp = sre_cuda_program_create(prog);
    if (cp == NULL) {
        return NULL;
    }
    if (!cp->has_dfa && !cp->has_nfa) {
        fail("program too large for the GPU Thompson tiers");
        return NULL;
    }
    sre_vm_thompson_ctx_t *ctx = static_cast<sre_vm_thompson_ctx_t *>(sre_pcalloc(pool, sizeof(*ctx)));
    if (ctx == NULL) {
        return NULL;
    }
    ctx->cp = cp;
    ctx->started = false;
    if (thompson_reserve(ctx, 4096) != SRE_OK
        || sre_pool_add_cleanup(pool, thompson_ctx_cleanup, ctx) != SRE_OK)
    {
        cudaFree(ctx->d_mem);
        return NULL;
    }
    return ctx;
}

/*
 * reference: sre_vm_thompson_exec, sre_vm_thompson.c:63-270.  Same contract:
 * SRE_OK as soon as a step of this call sees a live MATCH thread, SRE_AGAIN
 * when !eof, SRE_DECLINED at eof; state is carried between calls (on the
 * device; for the chunk-parallel path as one DFA state in the context).
 */
SRE_API sre_int_t
sre_vm_thompson_exec(sre_vm_thompson_ctx_t *ctx, sre_char *input, size_t size, unsigned eof)
{
    if (ctx == NULL || ctx->cp == NULL) {
        return SRE_ERROR;
    }
    sre_cuda_program_t *cp = ctx->cp;
    const cudaStream_t st = CLASSIC_STREAM;
    if (thompson_reserve(ctx, size) != SRE_OK) {
        return SRE_ERROR;
    }
    uint32_t *d_state = reinterpret_cast<uint32_t *>(ctx->d_mem);
    int32_t *d_rc = reinterpret_cast<int32_t *>(ctx->d_mem + 4096 * 4);
    uint8_t *d_in = ctx->d_mem + THOMPSON_HDR;
    if (size) {
        CUDA_TRY(cudaMemcpyAsync(d_in, input, size, cudaMemcpyHostToDevice, st));
    }

    /* long buffers over a determinised program: the chunk-parallel scan */
    if (cp->has_dfa && cp->has_image && size >= (1u << 16)) {
        uint32_t state = SRE_CUDA_STATE_INIT;
        if (ctx->started) {
            CUDA_TRY(cudaMemcpyAsync(&state, d_state, 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        const int rc = sre_cuda_thompson_exec_stream(cp, d_in, size, size, eof, &state, NULL, st);
        if (rc == SRE_ERROR) {
            return SRE_ERROR;
        }
        CUDA_TRY(cudaMemcpyAsync(d_state, &state, 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        ctx->started = true;
        return rc;
    }

    int launches = 0;
    cudaError_t err;
    if (cp->has_dfa) {
        err = sre_launch_dfa_carry(cp->dfa, d_in, nullptr, 1, 0, size, d_state, !ctx->started, eof != 0,
                                   d_rc, st, &launches);
    } else {
        err = sre_launch_nfa_lines(cp->nfa, d_in, nullptr, 1, 0, size, d_state, !ctx->started, eof != 0,
                                   d_rc, st, &launches);
    }
    count_launches(launches);
    if (err != cudaSuccess) {
        return fail("Thompson kernel launch failed: %s", cudaGetErrorString(err));
    }
    ctx->started = true;
    int32_t rc = SRE_ERROR;
    CUDA_TRY(cudaMemcpyAsync(&rc, d_rc, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return rc;
}

/* reference: sre_vm_thompson_jit_compile, sre_vm_thompson_jit.c:39-155.  Here:
 * lowering + determinisation + upload (done once, cached in the program). */
SRE_API sre_int_t
sre_vm_thompson_jit_compile(sre_pool_t *pool, sre_program_t *prog, sre_vm_thompson_code_t **pcode)
{
    *pcode = NULL;
    if (!device_ok()) {
        return SRE_DECLINED;    /* like the reference on a non-x86-64 target */
    }
    sre_cuda_program_t *cp = sre_cuda_program_create(prog);
    if (cp == NULL) {
        return SRE_ERROR;
    }
    sre_vm_thompson_code_t *code = static_cast<sre_vm_thompson_code_t *>(sre_pcalloc(pool, sizeof(*code)));
    if (code == NULL) {
        return SRE_ERROR;
    }
    code->cp = cp;
    *pcode = code;
    return SRE_OK;
}

SRE_API sre_vm_thompson_ctx_t *
sre_vm_thompson_jit_create_ctx(sre_pool_t *pool, sre_program_t *prog)
{
    return sre_vm_thompson_create_ctx(pool, prog);
}

SRE_API sre_vm_thompson_exec_pt
sre_vm_thompson_jit_get_handler(sre_vm_thompson_code_t *code)
{
    (void) code;
    return sre_vm_thompson_exec;
}

SRE_API sre_int_t
sre_vm_thompson_jit_free(sre_vm_thompson_code_t *code)
{
    (void) code;        /* tables belong to the program's pool */
    return SRE_OK;
}

/* reference: sre_vm_pike_create_ctx, sre_vm_pike.c:94-145 */
SRE_API sre_vm_pike_ctx_t *
sre_vm_pike_create_ctx(sre_pool_t *pool, sre_program_t *prog, sre_int_t *ovector, size_t ovecsize)
{
    sre_cuda_program_t *cp = sre_cuda_program_create(prog);
    if (cp == NULL) {
        return NULL;
    }
    sre_vm_pike_ctx_t *ctx = static_cast<sre_vm_pike_ctx_t *>(sre_pcalloc(pool, sizeof(*ctx)));
    if (ctx == NULL) {
        return NULL;
    }
    ctx->cp = cp;
    ctx->ovector = ovector;
    ctx->ovec_slots = ovecsize / sizeof(sre_int_t);
    ctx->h_out = new (std::nothrow) std::vector<int64_t>(4 + ctx->ovec_slots + 2, -1);
    ctx->d_in_cap = 4096;
    int launches = 0;
    if (ctx->h_out == NULL
        || cudaMalloc(&ctx->d_ctx, cp->pike.ctx_stride) != cudaSuccess
        || cudaMalloc(&ctx->d_out, (4 + ctx->ovec_slots + 2) * 8) != cudaSuccess
        || cudaMalloc(&ctx->d_in, ctx->d_in_cap) != cudaSuccess
        || sre_launch_pike_ctx_init(cp->pike, ctx->d_ctx, CLASSIC_STREAM, &launches) != cudaSuccess
        || cudaStreamSynchronize(CLASSIC_STREAM) != cudaSuccess
        || sre_pool_add_cleanup(pool, pike_ctx_cleanup, ctx) != SRE_OK)
    {
        fail("creating the Pike context failed: %s", cudaGetErrorString(cudaGetLastError()));
        pike_ctx_cleanup(ctx);
        return NULL;
    }
    count_launches(launches);
    return ctx;
}

/* reference: sre_vm_pike_exec, sre_vm_pike.c:148-689 */
SRE_API sre_int_t
sre_vm_pike_exec(sre_vm_pike_ctx_t *ctx, sre_char *input, size_t size, unsigned eof,
    sre_int_t **pending_matched)
{
    if (ctx == NULL || ctx->cp == NULL || ctx->d_ctx == NULL) {
        return SRE_ERROR;
    }
    if (size > ctx->d_in_cap) {
        cudaFree(ctx->d_in);
        ctx->d_in = nullptr;
        ctx->d_in_cap = (size + 4095) & ~(size_t) 4095;
        CUDA_TRY(cudaMalloc(&ctx->d_in, ctx->d_in_cap));
    }
    if (size) {
        CUDA_TRY(cudaMemcpyAsync(ctx->d_in, input, size, cudaMemcpyHostToDevice, CLASSIC_STREAM));
    }
    int launches = 0;
    size_t skip = 0;
    /*
     * One long buffer on a fresh context (bench/sregex.c:330-334): the Pike VM need not walk
     * what cannot be part of the match.  The chunk-parallel Thompson scan says whether and at
     * which step a match is first seen; the restart flags of the hint automaton give, from
     * there backwards, the last position before it at which no thread was alive; the Pike VM
     * runs from that position (as if the bytes before had been fed earlier) until its list is
     * empty.  Same rc and ovector: leftmost-first cannot begin before a position at which
     * every thread is dead.
     */
    sre_cuda_program_t *cp = ctx->cp;
    /* (not when the reference's prefilter could misfire: that replay needs the whole buffer) */
    if (!ctx->started && eof && size >= (1u << 16) && cp->has_dfa && cp->has_image && cp->dfa.hcls != nullptr
        && !cp->pike.quirk_possible) {
        sre_cuda_stream_scan_t *sc = nullptr;
        if (stream_reduce(cp, ctx->d_in, size, nullptr, cp->dfa.start, CLASSIC_STREAM, &sc) != SRE_OK) {
            return SRE_ERROR;
        }
        uint32_t exit_state = 0;
        int64_t off = -1;
        int r = stream_resolve(sc, cp->dfa.start, &exit_state, &off, nullptr);
        if (r == SRE_OK && exit_state != cp->dfa.acc && !cp->low.dfa.fin[exit_state]) {
            delete sc;
            ctx->started = true;
            return SRE_DECLINED;        /* no thread ever matches: the Pike VM would walk it all to say so */
        }
        long long p0 = 0;
        if (r == SRE_OK) {
            const size_t limit = exit_state == cp->dfa.acc && off >= 0 ? (size_t) off + 1 : size;
            cudaError_t e = sre_launch_dfa_stream_restart(cp->dfa, ctx->d_in, size, cp->dfa.start, limit, sc->ws,
                                                          reinterpret_cast<long long *>(sc->d_res), CLASSIC_STREAM,
                                                          &launches);
            if (e == cudaSuccess) {
                e = cudaMemcpyAsync(&p0, sc->d_res, sizeof(p0), cudaMemcpyDeviceToHost, CLASSIC_STREAM);
            }
            if (e == cudaSuccess) {
                e = cudaStreamSynchronize(CLASSIC_STREAM);
            }
            if (e != cudaSuccess) {
                r = fail("restart kernel failed: %s", cudaGetErrorString(e));
            }
        }
        delete sc;
        if (r != SRE_OK) {
            return SRE_ERROR;
        }
        skip = p0 > 0 && (size_t) p0 <= size ? (size_t) p0 : 0;
    }
    ctx->started = true;
    cudaError_t err = sre_launch_pike_stream(ctx->cp->pike, ctx->d_ctx, ctx->d_in, size, skip, eof != 0,
                                             pending_matched != NULL, ctx->d_out,
                                             (uint32_t) ctx->ovec_slots, CLASSIC_STREAM, &launches);
    count_launches(launches);
    if (err != cudaSuccess) {
        return fail("Pike kernel launch failed: %s", cudaGetErrorString(err));
    }
    std::vector<int64_t> &out = *ctx->h_out;
    CUDA_TRY(cudaMemcpyAsync(out.data(), ctx->d_out, (4 + ctx->ovec_slots) * 8, cudaMemcpyDeviceToHost,
                             CLASSIC_STREAM));
    CUDA_TRY(cudaStreamSynchronize(CLASSIC_STREAM));

    const sre_int_t rc = (sre_int_t) out[0];
    if (rc >= 0) {
        for (size_t i = 0; i < ctx->ovec_slots; i++) {
            ctx->ovector[i] = (sre_int_t) out[4 + i];
        }
    } else if (rc == SRE_AGAIN) {
        /* temp captures: only $& is reported (sre_vm_pike.c:692-735) */
        if (ctx->ovec_slots > 0) ctx->ovector[0] = (sre_int_t) out[4];
        if (ctx->ovec_slots > 1) ctx->ovector[1] = (sre_int_t) out[5];
        if (pending_matched) {
            if (out[1]) {
                ctx->pending[0] = (sre_int_t) out[2];
                ctx->pending[1] = (sre_int_t) out[3];
                *pending_matched = ctx->pending;
            } else {
                *pending_matched = NULL;
            }
        }
    }
    return rc;
}

}  /* extern "C" */
