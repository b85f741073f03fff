// Consumer side of the fused detections exchange (include/mydet.h "exchange protocol").
// The producer is stage E of postprocess_small_kernel: rows + count into every rank's buffer, and a local count of how
// often each image has been written.  PUBLISHING them -- the image's sequence number, with release semantics at system
// scope -- is the first thing the next kernel on the producer's stream does (exchange_publish_kernel, or the prologue of
// the fused consumer): at a kernel boundary the producer's stores are complete, so one fence + a handful of flag stores
// replace a system-scope drain at the end of every producer CTA.  The kernels here let a consumer on ANY rank order
// itself against the producers of all ranks on the device -- no host barrier, no stream dependency across processes:
//   exchange_wait_kernel     acquire-polls the local copy's sequence numbers until every image has been published
//                            once more than this rank has consumed so far; snapshots the counts;
//   exchange_release_kernel  publishes "this rank has consumed publication k" into the ack word of EVERY rank's copy,
//                            which is what the producers' back-pressure wait reads (locally).
#include "common.cuh"
#include "internal.cuh"

namespace mydet {

struct ReleaseParams { unsigned* peer[8]; unsigned* mc; int n_peers, self, rank_slot; long long images_total; int out_cap, n_param;
                       long long own_row0; int own_rows; };

// seq[i] = prod[i] for this rank's images, in every rank's copy.  Runs after the producing kernel on its stream.
__device__ __forceinline__ void publish_rows(const ReleaseParams& R, const ExchangeLayout& XL) {
    __threadfence_system();      // the producing kernel's stores (complete at the kernel boundary) before the flags, for every observer
    for (int i = threadIdx.x; i < R.own_rows; i += blockDim.x) {
        const long long row = R.own_row0 + i;
        const unsigned k = R.peer[R.self][XL.prod_off + row];
        if (R.mc) multimem_st_release_u32(R.mc + XL.seq_off + row, k);
        else for (int q = 0; q < R.n_peers; ++q) st_release_sys(R.peer[q] + XL.seq_off + row, k);
    }
}
__global__ void __launch_bounds__(256) exchange_publish_kernel(const ReleaseParams R) {
    publish_rows(R, exchange_layout(R.images_total, R.out_cap, R.n_param));
}

__device__ __forceinline__ void publish_ack(const ReleaseParams& R, const ExchangeLayout& XL, unsigned k) {
    // reads of the consumed rows (earlier kernels of this stream, or this CTA before its barrier) are complete; the fence
    // orders them before the ack becomes visible to the producers, which then overwrite the rows
    __threadfence_system();
    if (R.mc) multimem_st_release_u32(R.mc + XL.ack_off + R.rank_slot, k);
    else for (int q = 0; q < R.n_peers; ++q) st_release_sys(R.peer[q] + XL.ack_off + R.rank_slot, k);
}

// RELEASE: wait + snapshot + acknowledgement in one launch, for consumers that need nothing but the counts (or that
// copy what they need in the same kernel later on): the ack goes out as soon as every image has been seen.
template <bool RELEASE>
__global__ void __launch_bounds__(256) exchange_wait_kernel(unsigned* local, long long images_total, int out_cap, int n_param,
                                                            int* counts_out, int* status, const ReleaseParams R) {
    const ExchangeLayout XL = exchange_layout(images_total, out_cap, n_param);
    if (RELEASE && R.own_rows > 0) publish_rows(R, XL);                   // the fused consumer publishes its own rank's step first
    const unsigned want = ld_relaxed_sys(local + XL.want_off) + 1u;       // only this kernel / the release kernel write it
    const long long t0 = clock64();
    bool late = false;
    for (long long i = threadIdx.x; i < images_total; i += blockDim.x) {
        while ((int)(ld_acquire_sys(local + XL.seq_off + i) - want) < 0) {
            if (clock64() - t0 > kExchangeSpinCycles) { late = true; break; }
            __nanosleep(128);
        }
        // the acquire above orders this read (and every read of the image's rows by later kernels of the stream) behind
        // the producer's row / count stores
        if (counts_out) counts_out[i] = (int)ld_relaxed_sys(local + XL.counts_off + i);
    }
    const int any_late = __syncthreads_or(late ? 1 : 0);
    if (threadIdx.x == 0) {
        local[XL.want_off + 1] = want;                                     // staged: the release kernel commits it
        if (status) *status = any_late ? 1 : 0;
        if (RELEASE) { local[XL.want_off] = want; publish_ack(R, XL, want); }
    }
}

__global__ void exchange_release_kernel(const ReleaseParams R) {
    if (threadIdx.x != 0) return;
    const ExchangeLayout XL = exchange_layout(R.images_total, R.out_cap, R.n_param);
    unsigned* own = R.peer[R.self];
    const unsigned k = own[XL.want_off + 1];
    own[XL.want_off] = k;
    publish_ack(R, XL, k);
}

}  // namespace mydet

using namespace mydet;

MYDET_API size_t mydet_exchange_buffer_bytes(int64_t images_total, int out_cap, int n_param) {
    if (images_total < 0 || out_cap <= 0 || (n_param != 4 && n_param != 5)) return 0;
    return (size_t)exchange_layout(images_total, out_cap, n_param).total_words * 4;
}

MYDET_API int mydet_exchange_wait(void* local_buf, int64_t images_total, int out_cap, int n_param, int32_t* counts_out,
                                  int32_t* status, void* stream) {
    MYDET_REQUIRE(local_buf && images_total > 0 && out_cap > 0 && (n_param == 4 || n_param == 5), "bad exchange buffer description");
    MYDET_REQUIRE((reinterpret_cast<uintptr_t>(local_buf) & 15) == 0, "exchange buffer must be 16-byte aligned");
    exchange_wait_kernel<false><<<1, 256, 0, (cudaStream_t)stream>>>(static_cast<unsigned*>(local_buf), images_total, out_cap,
                                                                    n_param, counts_out, status, ReleaseParams{});
    return launch_status("exchange_wait_kernel");
}

static int fill_release(ReleaseParams& R, void* const* peer_bufs, int n_peers, void* multicast_buf, int self_index,
                        int64_t images_total, int out_cap, int n_param, int64_t own_row0 = 0, int own_rows = 0) {
    MYDET_REQUIRE(peer_bufs && n_peers >= 1 && n_peers <= 8 && self_index >= 0 && self_index < n_peers, "bad peer description");
    MYDET_REQUIRE(images_total > 0 && out_cap > 0 && (n_param == 4 || n_param == 5), "bad exchange buffer description");
    for (int q = 0; q < 8; ++q) R.peer[q] = q < n_peers ? static_cast<unsigned*>(peer_bufs[q]) : nullptr;
    R.mc = static_cast<unsigned*>(multicast_buf); R.n_peers = n_peers; R.self = self_index; R.rank_slot = self_index;
    R.images_total = images_total; R.out_cap = out_cap; R.n_param = n_param;
    MYDET_REQUIRE(own_rows >= 0 && own_row0 >= 0 && own_row0 + own_rows <= images_total, "bad image range of this rank");
    R.own_row0 = own_row0; R.own_rows = own_rows;
    return 0;
}

MYDET_API int mydet_exchange_publish(void* const* peer_bufs, int n_peers, void* multicast_buf, int self_index, int64_t image_offset,
                                     int batch, int64_t images_total, int out_cap, int n_param, void* stream) {
    ReleaseParams R;
    if (int rc = fill_release(R, peer_bufs, n_peers, multicast_buf, self_index, images_total, out_cap, n_param, image_offset, batch)) return rc;
    if (batch == 0) return 0;
    exchange_publish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(R);
    return launch_status("exchange_publish_kernel");
}

MYDET_API int mydet_exchange_consume_counts(void* const* peer_bufs, int n_peers, void* multicast_buf, int self_index,
                                            int64_t image_offset, int batch, int64_t images_total, int out_cap, int n_param,
                                            int32_t* counts_out, int32_t* status, void* stream) {
    ReleaseParams R;
    if (int rc = fill_release(R, peer_bufs, n_peers, multicast_buf, self_index, images_total, out_cap, n_param, image_offset, batch)) return rc;
    exchange_wait_kernel<true><<<1, 256, 0, (cudaStream_t)stream>>>(R.peer[self_index], images_total, out_cap, n_param,
                                                                   counts_out, status, R);
    return launch_status("exchange_wait_kernel<release>");
}

MYDET_API int mydet_exchange_release(void* const* peer_bufs, int n_peers, void* multicast_buf, int self_index, int64_t images_total,
                                     int out_cap, int n_param, void* stream) {
    ReleaseParams R;
    if (int rc = fill_release(R, peer_bufs, n_peers, multicast_buf, self_index, images_total, out_cap, n_param)) return rc;
    exchange_release_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(R);
    return launch_status("exchange_release_kernel");
}
