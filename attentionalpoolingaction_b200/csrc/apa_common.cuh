// apa_common.cuh — sm_100a PTX helpers shared by the attentional-pooling kernels.
// mbarrier / bulk-copy (TMA engine) / cluster + DSMEM / grid barrier.  No CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace apa {

#ifndef APA_SPIN_TIMEOUT_CYCLES
#define APA_SPIN_TIMEOUT_CYCLES (6000000000ll)  // ~3 s at 2 GHz: trap instead of hanging the GPU
#endif
#ifndef APA_PEER_TIMEOUT_CYCLES
#define APA_PEER_TIMEOUT_CYCLES (120000000000ll)  // ~60 s: a straggling RANK (first-step skew, data stall) is not a bug
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Arrive on the same-named barrier of CTA `cta` of this cluster (release at cluster scope).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* local_bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > APA_SPIN_TIMEOUT_CYCLES) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > APA_SPIN_TIMEOUT_CYCLES) __trap();
  }
}

// ----------------------------------------------------------------------------- bulk copy (TMA engine, UBLKCP)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> this CTA's shared memory, completion counted in bytes on `bar`.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ----------------------------------------------------------------------------- cluster / DSMEM
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() {
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_dsmem_f4(uint32_t raddr, const float4& v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_dsmem_f(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
// volatile (keeps its order w.r.t. the barrier asms) but no memory clobber, so a run of these pipelines
__device__ __forceinline__ float ld_dsmem_f_nv(uint32_t raddr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr));
  return v;
}
// All threads of all CTAs of the cluster.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_ptr, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_ptr)), "r"(cta));
  return raddr;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t raddr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(raddr)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_dsmem_f(uint32_t raddr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
  return v;
}

// ----------------------------------------------------------------------------- programmatic dependent launch
// launch_dependents: the next kernel in the stream (if launched with the programmatic-serialization
// attribute) may start once every CTA of this grid has executed this (or exited).
// wait: block until the previous grid in the stream has fully completed and its writes are visible
// (no-op when this grid was not launched as a programmatic dependent).
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ----------------------------------------------------------------------------- fused all-gather (peer stores)
// flags[q] is rank q's flag array of 2 * world words: word r = the last step whose logits rank r has completely
// stored into q's gathered ring; word world + r = the last step rank r has STARTED after everything before it in
// r's stream completed (= r has consumed every gathered buffer it had to, see include/apa.h).
// ONE system-scope fence, then a relaxed store per peer (a release store per peer would be a fence per peer, i.e. one
// NVLink round trip after the other)
__device__ __forceinline__ void peer_gather_publish(int world, int rank, unsigned long long* const* flags,
                                                    unsigned long long step) {
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  for (int q = 0; q < world; ++q)
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(flags[q] + rank), "l"(step) : "memory");
}
// "started" releases nothing (it says that kernels which ended before this one no longer read the ring): no fence
__device__ __forceinline__ void peer_gather_publish_started(int world, int rank, unsigned long long* const* flags,
                                                            unsigned long long step) {
  for (int q = 0; q < world; ++q)
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(flags[q] + world + rank), "l"(step) : "memory");
}
// one thread: spin until word `first + r` of my flag array is >= target for every rank r
__device__ __forceinline__ void peer_gather_wait_all(int world, const unsigned long long* mine, int first,
                                                     unsigned long long target) {
  const long long t0 = clock64();
  for (int r = 0; r < world; ++r) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + first + r) : "memory");
      if (v >= target) break;
      __nanosleep(100);
      if (clock64() - t0 > APA_PEER_TIMEOUT_CYCLES) __trap();
    } while (true);
  }
}

// The step of the gather call a kernel belongs to: calls completed so far + 1.  Only valid after the kernel's
// griddepcontrol.wait (the previous call, which bumps the count at its very end, has then completed).
__device__ __forceinline__ unsigned long long peer_gather_step(const unsigned long long* counters) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counters + 1) : "memory");
  return v + 1;
}
// The launch's very last act (one thread, after every logits store of the launch is ordered before it).
__device__ __forceinline__ void peer_gather_complete(int world, int rank, unsigned long long* const* flags,
                                                     unsigned long long* counters, unsigned long long step) {
  peer_gather_publish(world, rank, flags, step);
  asm volatile("red.release.gpu.global.max.u64 [%0], %1;" ::"l"(counters + 1), "l"(step) : "memory");
}
// Older forms (grid-barrier / cluster-local / pipeline): called by ONE thread per CTA after that CTA's logits
// stores (and a CTA-level barrier).  counters[0] only ever grows and every launch leaves it at a multiple of 2^20,
// so the grid may change between calls; the last CTA of the launch to get here completes the step.
__device__ __forceinline__ void peer_gather_finish(int world, int rank, unsigned long long* const* flags,
                                                   unsigned long long* counters, unsigned int nblocks,
                                                   unsigned long long step) {
  if (world > 1) __threadfence_system();  // this CTA's peer stores are globally visible before it is counted
  unsigned long long prev;
  asm volatile("atom.add.acq_rel.gpu.global.u64 %0, [%1], 1;" : "=l"(prev) : "l"(counters) : "memory");
  if ((prev & ((1ull << 20) - 1)) != (unsigned long long)(nblocks - 1)) return;
  const unsigned long long pad = (1ull << 20) - nblocks;
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(counters), "l"(pad) : "memory");
  peer_gather_complete(world, rank, flags, counters, step);
}

// ----------------------------------------------------------------------------- misc
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float4 ldg_cg_f4(const float* p) {  // L2-coherent (bypass L1)
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float ldg_cg_f(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Grid-wide barrier for a grid whose CTAs are all co-resident (grid <= resident capacity; the host
// launches cooperatively when the driver allows it).  `ctr` is a 64-bit arrival counter that only ever
// grows: every launch with the same grid size adds exactly `nblocks`, so the next multiple of `nblocks`
// above a CTA's own ticket is the release point.  One counter per grid size (the host indexes the
// workspace control block by gridDim), zeroed once by apa_attn_pool_workspace_init.
// Split in two so independent work can be done between arrive and wait; call both from ONE thread,
// with a __syncthreads() before the arrive and after the wait.
__device__ __forceinline__ unsigned long long grid_arrive(unsigned long long* ctr, unsigned int nblocks) {
  unsigned long long prev;
  asm volatile("atom.add.release.gpu.global.u64 %0, [%1], 1;" : "=l"(prev) : "l"(ctr) : "memory");
  return prev - (prev % nblocks) + nblocks;
}
__device__ __forceinline__ void grid_wait(unsigned long long* ctr, unsigned long long target) {
  const long long t0 = clock64();
  unsigned long long v;
  do {
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    if (clock64() - t0 > APA_SPIN_TIMEOUT_CYCLES) __trap();
  } while (true);
}

}  // namespace apa
