// SyncBN statistics exchange over NVLink peer memory (replaces the SyncMaster / SlavePipe queue rendezvous and the
// ReduceAddCoalesced + Broadcast pair of batchnorm.py:95-112, comm.py:74-136).
//
// One single-CTA kernel per exchange: every rank stores its [sum | sum of squares] vector straight into EVERY peer's
// receive slot (plain st.global over NVLink / NVSwitch; the buffers are symmetric allocations mapped into all ranks'
// address spaces), raises a per-source flag at the peer with a system-scope release store, spins on its own flags with
// system-scope acquire loads, then adds the world's vectors in rank order -- all ranks therefore compute bit-identical
// statistics.  Messages are 1-12 KB, so the exchange is pure latency: one launch + one NVLink store/flag round trip
// instead of NCCL's ~20 us small-message all-reduce (bench.py: syncbn_stat_reduce_us).
//
// Buffer layout on every rank (doubles):  data [2 slots][world][slot_doubles]  then  flags (uint32) [2 slots][world].
// Slots alternate with the call counter (epoch), kept in device memory so the launch carries no per-call host scalar
// and can sit in a CUDA graph.  Two slots suffice: a rank cannot finish exchange e + 1 before every peer has raised
// its e + 1 flag, which a peer does only after it has finished reading slot e.
#include "common.cuh"

namespace ssg {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(512) p2p_allreduce_f64_kernel(double* __restrict__ data, int n, void* const* __restrict__ peers, int rank,
                                                                 int world, int slot_doubles, unsigned* __restrict__ epoch,
                                                                 long long timeout_ns, unsigned* __restrict__ status) {
    __shared__ unsigned e_s;
    if (threadIdx.x == 0) e_s = *epoch + 1u;
    __syncthreads();
    const unsigned e = e_s;
    const int slot = (int)(e & 1u);
    const size_t flags_off = (size_t)2 * world * slot_doubles;                  // in doubles
    // 1. push the local vector into slot [slot][rank] of every rank (self included: keeps the reduction loop uniform)
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(peers[p]) + ((size_t)slot * world + rank) * slot_doubles;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = data[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise this rank's flag at every peer; 3. wait until every source has raised its flag here
    if ((int)threadIdx.x < world) {
        unsigned* pf = reinterpret_cast<unsigned*>(reinterpret_cast<double*>(peers[threadIdx.x]) + flags_off);
        st_release_sys(pf + slot * world + rank, e);
        const unsigned* mf = reinterpret_cast<const unsigned*>(reinterpret_cast<const double*>(peers[rank]) + flags_off) + slot * world + threadIdx.x;
        // A peer that never arrives (rank skew beyond the bound, a dead rank) must not kill the context: after timeout_ns of
        // wall clock the wait is abandoned, the failure is recorded in *status and the host raises when it next checks.
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire_sys(mf) != e) {
            if ((++spins & 0x3ffu) == 0 && timeout_ns > 0) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > (unsigned long long)timeout_ns) {
                    if (status) atomicCAS(status, 0u, 1u + threadIdx.x);
                    break;
                }
            }
        }
    }
    __syncthreads();
    // 4. reduce in rank order (remote-written lines: volatile loads, never a stale L1 copy)
    const volatile double* mine = reinterpret_cast<const volatile double*>(peers[rank]) + (size_t)slot * world * slot_doubles;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double a = 0.0;
        for (int q = 0; q < world; ++q) a += mine[(size_t)q * slot_doubles + i];
        data[i] = a;
    }
    if (threadIdx.x == 0) *epoch = e;
}

}  // namespace ssg
using namespace ssg;

extern "C" {

long long ssg_p2p_buffer_bytes(int world, int slot_doubles) {
    return (long long)2 * world * slot_doubles * 8 + (long long)2 * world * 4 + 256;
}

int ssg_p2p_allreduce_f64(double* data, int n, void* const* peer_bufs_dev, int rank, int world, int slot_doubles, unsigned* epoch_dev,
                          ssg_stream_t s) {
    SSG_CHECK_ARG(data && peer_bufs_dev && epoch_dev, "p2p_allreduce: null pointer");
    SSG_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world, "p2p_allreduce: rank %d of %d", rank, world);
    SSG_CHECK_ARG(n > 0 && n <= slot_doubles, "p2p_allreduce: %d doubles do not fit a %d-double slot", n, slot_doubles);
    p2p_allreduce_f64_kernel<<<1, 512, 0, (cudaStream_t)s>>>(data, n, peer_bufs_dev, rank, world, slot_doubles, epoch_dev, 0, nullptr);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_p2p_allreduce_f64_to(double* data, int n, void* const* peer_bufs_dev, int rank, int world, int slot_doubles, unsigned* epoch_dev,
                             long long timeout_ns, unsigned* status_dev, ssg_stream_t s) {
    SSG_CHECK_ARG(data && peer_bufs_dev && epoch_dev, "p2p_allreduce: null pointer");
    SSG_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world, "p2p_allreduce: rank %d of %d", rank, world);
    SSG_CHECK_ARG(n > 0 && n <= slot_doubles, "p2p_allreduce: %d doubles do not fit a %d-double slot", n, slot_doubles);
    p2p_allreduce_f64_kernel<<<1, 512, 0, (cudaStream_t)s>>>(data, n, peer_bufs_dev, rank, world, slot_doubles, epoch_dev, timeout_ns,
                                                             status_dev);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
