/*
 * sre_lines.cu -- line index of a '\n'-delimited buffer, on the device.
 *
 * New in this build (no reference counterpart: the reference is handed one
 * buffer per sre_vm_*_exec call by its caller, e.g. bench/sregex.c:193-370
 * reads a whole file).  The batch entry points take either fixed-pitch lines
 * or an offsets array; this produces that array for text as it comes from a
 * log file, so that sre_cuda_thompson_exec_ragged / sre_cuda_pike_exec_lines
 * can run on it without a pass over the data on the host:
 *
 *   offsets[0] = 0, offsets[i + 1] = one past the '\n' that ends line i; a last
 *   line without '\n' ends at len.  Line i = buf[offsets[i], offsets[i + 1]),
 *   terminator included (what getline() returns; `$` still matches before it,
 *   sre_vm_thompson.c:183-188).
 *
 * Three kernels: newlines per 16 KB block; exclusive scan of the block counts
 * (one block); per block, ranks by warp / block prefix sums and the stores.
 * HBM-bound: the buffer is read twice (the second pass mostly from L2 for
 * buffers up to its size), 8 bytes written per line.
 */
#include "sre_kernels.cuh"

namespace {

constexpr uint32_t LB_THREADS = 256;
constexpr uint32_t LB_BYTES = 16384;        /* bytes per block */
constexpr uint32_t LB_PER_THREAD = LB_BYTES / LB_THREADS;   /* 64: four 16-byte loads */

/* bit 7 of every byte of x that equals '\n' */
__device__ __forceinline__ uint32_t nl_mask(uint32_t x)
{
    const uint32_t y = x ^ 0x0a0a0a0au;
    /* exact zero-byte test (no carry between bytes) */
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}

/* newlines among the LB_PER_THREAD bytes of this thread; bytes past len do not count */
__device__ __forceinline__ uint32_t thread_count(const uint8_t *buf, size_t begin, size_t len)
{
    uint32_t n = 0;
    if (begin + LB_PER_THREAD <= len && ((reinterpret_cast<uintptr_t>(buf) + begin) & 15) == 0) {
#pragma unroll
        for (uint32_t k = 0; k < LB_PER_THREAD; k += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(buf + begin + k));
            n += __popc(nl_mask(v.x)) + __popc(nl_mask(v.y)) + __popc(nl_mask(v.z)) + __popc(nl_mask(v.w));
        }
    } else {
        for (size_t p = begin; p < begin + LB_PER_THREAD && p < len; p++) {
            n += buf[p] == '\n';
        }
    }
    return n;
}

/* block-wide exclusive prefix sum of v; *total = sum over the block */
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[LB_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) {
            incl += up;
        }
    }
    if (lane == 31) {
        warp_sums[warp] = incl;
    }
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (uint32_t w = 0; w < LB_THREADS / 32; w++) {
        const uint32_t s = warp_sums[w];
        before += w < warp ? s : 0;
        all += s;
    }
    __syncthreads();
    *total = all;
    return before + incl - v;
}

__global__ void __launch_bounds__(LB_THREADS)
k_lines_count(const uint8_t *__restrict__ buf, size_t len, unsigned long long *__restrict__ counts)
{
    const size_t begin = (size_t) blockIdx.x * LB_BYTES + (size_t) threadIdx.x * LB_PER_THREAD;
    uint32_t total;
    block_exclusive(thread_count(buf, begin, len), &total);
    if (threadIdx.x == 0) {
        counts[blockIdx.x] = total;
    }
}

/* counts[i] <- sum of counts[0 .. i); counts[n] <- the number of lines: the
 * newlines, plus one for a last line that no newline ends (its end offset is
 * stored here).  One block. */
__global__ void __launch_bounds__(1024)
k_lines_scan(unsigned long long *counts, size_t n, const uint8_t *__restrict__ buf, size_t len,
             int64_t *__restrict__ offsets, size_t max_lines)
{
    __shared__ unsigned long long carry, sums[32];
    if (threadIdx.x == 0) {
        carry = 0;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t base = 0; base < n; base += 1024) {
        const size_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? counts[i] : 0;
        unsigned long long incl = v;
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) {
                incl += up;
            }
        }
        if (lane == 31) {
            sums[warp] = incl;
        }
        __syncthreads();
        unsigned long long before = 0, all = 0;
        for (uint32_t w = 0; w < 32; w++) {
            before += w < warp ? sums[w] : 0;
            all += sums[w];
        }
        const unsigned long long c = carry;
        if (i < n) {
            counts[i] = c + before + incl - v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            carry = c + all;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long lines = carry;
        if (len > 0 && buf[len - 1] != '\n') {
            if (lines < max_lines) {
                offsets[lines + 1] = (int64_t) len;
            }
            lines++;
        }
        counts[n] = lines;
    }
}

__global__ void __launch_bounds__(LB_THREADS)
k_lines_write(const uint8_t *__restrict__ buf, size_t len, const unsigned long long *__restrict__ counts,
              int64_t *__restrict__ offsets, size_t max_lines)
{
    const size_t begin = (size_t) blockIdx.x * LB_BYTES + (size_t) threadIdx.x * LB_PER_THREAD;
    uint32_t total;
    const uint32_t before = block_exclusive(thread_count(buf, begin, len), &total);
    size_t rank = (size_t) counts[blockIdx.x] + before;       /* newlines before this thread's bytes */
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        offsets[0] = 0;
    }
    if (begin + LB_PER_THREAD <= len && ((reinterpret_cast<uintptr_t>(buf) + begin) & 15) == 0) {
        /* 16-byte loads; the set bits of the equality mask give the positions */
#pragma unroll
        for (uint32_t k = 0; k < LB_PER_THREAD; k += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(buf + begin + k));
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                uint32_t m = nl_mask(w[j]);
                while (m) {
                    const uint32_t byte = (uint32_t) (__ffs((int) m) - 1) >> 3;
                    m &= m - 1;
                    if (rank < max_lines) {
                        offsets[rank + 1] = (int64_t) (begin + k + 4 * j + byte + 1);
                    }
                    rank++;
                }
            }
        }
        return;
    }
    for (size_t p = begin; p < begin + LB_PER_THREAD && p < len; p++) {
        if (buf[p] == '\n') {
            if (rank < max_lines) {
                offsets[rank + 1] = (int64_t) (p + 1);
            }
            rank++;
        }
    }
}

}  // namespace

size_t sre_lines_workspace_bytes(size_t len)
{
    const size_t nblocks = (len + LB_BYTES - 1) / LB_BYTES;
    return (nblocks + 1) * sizeof(unsigned long long);
}

cudaError_t sre_launch_index_lines(const uint8_t *buf, size_t len, int64_t *offsets, size_t max_lines,
    unsigned long long *workspace, cudaStream_t stream, int *launches)
{
    const size_t nblocks = (len + LB_BYTES - 1) / LB_BYTES;
    if (nblocks == 0) {             /* empty buffer: no line; offsets[0] = 0 */
        cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(unsigned long long), stream);
        return e != cudaSuccess ? e : cudaMemsetAsync(offsets, 0, sizeof(int64_t), stream);
    }
    if (nblocks > 0x7fffffffull) {
        return cudaErrorInvalidValue;
    }
    if (launches) {
        *launches += 3;
    }
    k_lines_count<<<(unsigned) nblocks, LB_THREADS, 0, stream>>>(buf, len, workspace);
    k_lines_scan<<<1, 1024, 0, stream>>>(workspace, nblocks, buf, len, offsets, max_lines);
    k_lines_write<<<(unsigned) nblocks, LB_THREADS, 0, stream>>>(buf, len, workspace, offsets, max_lines);
    return cudaGetLastError();
}
