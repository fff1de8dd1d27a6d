/*
 * sre_pike_lineage.cu -- Pike VM with submatch captures over the determinised
 * thread lists (../lower/sre_pdfa.h), one CUDA thread per line.
 *
 * What it replaces: sre_vm_pike_exec + sre_vm_pike_add_thread + the capture
 * copies (reference sre_vm_pike.c:148-689, :756-942, sre_capture.c) for
 * programs without assertions.  Instead of simulating the thread list (k_pike_table:
 * ~500 instructions per byte, 10 of 32 lanes busy), a lane runs
 *
 *   forward   one look-up per byte in the P-DFA (state = the ordered thread
 *             list), storing the state of each position in a 64-entry ring in
 *             shared memory and noting the last step that reported a match;
 *             it ends when the list is empty (every thread of higher priority
 *             than the match has died) or the line does;
 *   backward  from the thread that matched, parent by parent through the
 *             transitions' provenance records, giving each capture slot the
 *             position of the last step that SAVEd it -- until the lineage
 *             reaches the ".*?" thread or the start closure.
 *
 * Every lane executes the same two loops: no per-lane thread lists, no
 * divergence beyond the span lengths.  A lineage older than the ring (a match
 * longer than ~60 bytes) is reported SRE_K_RETRY and re-run by k_pike_table.
 */
#include "sre_device_common.cuh"

using namespace sre_dev;

namespace {

constexpr int TB = 128;
constexpr int RING = 64;            /* positions remembered per lane (power of two) */

template <bool SMEM_TAB>
__global__ void __launch_bounds__(TB)
k_pike_lineage(sre_dev_pdfa_t d, const uint32_t *__restrict__ slot_ofs, const uint8_t *__restrict__ buf,
               const int64_t *__restrict__ offsets, size_t nlines, size_t pitch, size_t linelen,
               sre_line_list_t lines, const int32_t *__restrict__ start_hint, int32_t *__restrict__ rc,
               int64_t *__restrict__ ovec, uint32_t ovec_slots, sre_pike_work_t *work)
{
    extern __shared__ __align__(16) uint8_t smem[];
    /* [ring: RING x TB u16][clsmap 256][trans (SMEM_TAB)] */
    uint16_t *ring = reinterpret_cast<uint16_t *>(smem) + threadIdx.x;
    uint8_t *s_cls = smem + RING * TB * 2;
    const uint16_t *trans = d.trans;
    for (uint32_t i = threadIdx.x; i < 256; i += TB) {
        s_cls[i] = d.clsmap[i];
    }
    if (SMEM_TAB) {
        uint16_t *s_trans = reinterpret_cast<uint16_t *>(s_cls + 256);
        for (uint32_t i = threadIdx.x; i < d.nstates * d.nclasses; i += TB) {
            s_trans[i] = d.trans[i];
        }
        trans = s_trans;
    }
    __syncthreads();
    const uint32_t C = d.nclasses;
    const size_t nwork = lines.list ? (size_t) *lines.count : nlines;

    for (size_t k = (size_t) blockIdx.x * TB + threadIdx.x; k < nwork; k += (size_t) gridDim.x * TB) {
        const size_t line = lines.list ? (size_t) lines.list[k] : k;
        const size_t lstart = offsets ? (size_t) offsets[line] : line * pitch;
        const size_t lend = offsets ? (size_t) offsets[line + 1] : lstart + linelen;
        const uint8_t *input = buf + lstart;
        const int32_t size = (int32_t) (lend - lstart);
        const int32_t start = start_hint ? start_hint[line] : 0;
        int64_t *ov = ovec + line * ovec_slots;

        /* forward */
        uint32_t s = d.init;
        int32_t pos = start, mpos = -1;
        while (pos < size && s != 0) {
            /* the aligned 16 bytes that hold input[pos] */
            const uintptr_t at = reinterpret_cast<uintptr_t>(input + pos);
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(at & ~(uintptr_t) 15));
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
            int32_t q = (int32_t) (at & 15);
            const int32_t qend = size - pos < 16 - q ? q + (size - pos) : 16;
#pragma unroll 1
            for (; q < qend && s != 0; q++, pos++) {
                const uint32_t b = (w[q >> 2] >> ((q & 3) * 8)) & 0xff;
                ring[(pos & (RING - 1)) * TB] = (uint16_t) s;
                const uint32_t e = SMEM_TAB ? trans[s * C + s_cls[b]] : __ldg(trans + s * C + s_cls[b]);
                mpos = (e & 0x8000u) ? pos : mpos;
                s = e & 0x7fffu;
            }
        }
        const bool at_eof = s != 0 && __ldg(d.eof_idx + s) != 0xff;
        for (uint32_t i = 0; i < ovec_slots; i++) {
            ov[i] = -1;
        }
        if (mpos < 0 && !at_eof) {
            rc[line] = SRE_K_DECLINED;
            continue;
        }

        /* backward: j indexes the list of state `cur`, the state before step u + 1 */
        const int32_t oldest = pos - RING;      /* positions > oldest are still in the ring */
        uint32_t cur, j, rid, unset = 0xffffffffu;
        int32_t u;
        bool lost = false;
        auto assign = [&](uint32_t mask, int32_t where) {
            uint32_t m = mask & unset;
            unset &= ~mask;
            while (m) {
                const uint32_t slot = __ffs(m) - 1;
                m &= m - 1;
                if (slot < ovec_slots) {
                    ov[slot] = where;
                }
            }
        };
        if (at_eof) {
            cur = s;
            j = __ldg(d.eof_idx + s);
            rid = __ldg(d.eof_regex + s);
            u = size - 1;
        } else if (mpos <= oldest) {
            lost = true;
            cur = j = rid = 0;
            u = -1;
        } else {
            cur = ring[(mpos & (RING - 1)) * TB];
            const uint32_t t = cur * C + s_cls[input[mpos]];
            j = __ldg(d.mparent + t);
            rid = __ldg(d.mregex + t);
            assign(__ldg(d.mmask + t), mpos + 1);
            u = mpos - 1;
        }
        while (!lost) {
            if (j == __ldg(d.any_idx + cur)) {
                break;                      /* the ".*?" thread carries no captures */
            }
            if (u < start) {
                assign(__ldg(d.init_mask + j), start);      /* a thread of the start closure */
                break;
            }
            if (u <= oldest) {
                lost = true;
                break;
            }
            const uint32_t before = ring[(u & (RING - 1)) * TB];
            const uint32_t idx = __ldg(d.eofs + before * C + s_cls[input[u]]) + j;
            assign(__ldg(d.emask + idx), u + 1);
            j = __ldg(d.eparent + idx);
            cur = before;
            u--;
        }
        if (lost) {
            /* the lineage is older than the ring: the next tier re-runs the line */
            rc[line] = SRE_K_RETRY;
            atomicAdd(&work->given_up[0], 1u);
            continue;
        }
        rc[line] = (int32_t) rid;
        (void) slot_ofs;
    }
}

}  // namespace

bool sre_pike_lineage_applicable(const sre_dev_pdfa_t &d, size_t linelen)
{
    return d.nstates != 0 && linelen < (1ull << 31);
}

/* same contract as sre_launch_pike_table, pass 0: lines it gives up on get SRE_K_RETRY and are
 * counted in work->given_up[0] */
cudaError_t sre_launch_pike_lineage(const sre_dev_pdfa_t &d, const sre_dev_pike_t &pk, const uint8_t *buf,
    const int64_t *offsets, size_t nlines, size_t pitch, size_t linelen, sre_line_list_t lines,
    const int32_t *start, int32_t *rc, int64_t *ovec, uint32_t ovec_slots, sre_pike_work_t *work,
    cudaStream_t stream, int *launches)
{
    if (nlines == 0) {
        return cudaSuccess;
    }
    if (launches) {
        ++*launches;
    }
    cudaError_t ce = cudaMemsetAsync(work, 0, sizeof(*work), stream);
    if (ce != cudaSuccess) {
        return ce;
    }
    const size_t tab = (size_t) d.nstates * d.nclasses * 2;
    const bool smem_tab = tab <= 64 * 1024;
    const size_t smem = (size_t) RING * TB * 2 + 256 + (smem_tab ? tab : 0);
    size_t grid = (nlines + TB - 1) / TB;
    const size_t cap = (size_t) num_sms() * (smem_tab && tab > 8 * 1024 ? 2 : 8);
    if (grid > cap) {
        grid = cap;
    }
    if (smem_tab) {
        static size_t smem_set = 0;
        if (smem > smem_set) {
            ce = cudaFuncSetAttribute(k_pike_lineage<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (ce != cudaSuccess) {
                return ce;
            }
            smem_set = smem;
        }
        k_pike_lineage<true><<<(unsigned) grid, TB, smem, stream>>>(d, pk.slot_ofs, buf, offsets, nlines, pitch,
                                                                    linelen, lines, start, rc, ovec, ovec_slots, work);
    } else {
        k_pike_lineage<false><<<(unsigned) grid, TB, smem, stream>>>(d, pk.slot_ofs, buf, offsets, nlines, pitch,
                                                                     linelen, lines, start, rc, ovec, ovec_slots, work);
    }
    return cudaGetLastError();
}
