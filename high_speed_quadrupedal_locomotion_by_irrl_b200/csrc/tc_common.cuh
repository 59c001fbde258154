// tcgen05 / TMEM / bulk-copy / mbarrier building blocks shared by the tensor-memory kernels (policy_tc_kernels.cu: LSTM act;
// learner_tc.cu: the learner's streaming projections).  sm_100a only.
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

namespace irrl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch failure reported to the caller) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    if (mbar_try_wait(b, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(b, parity)) { if (clock64() - t0 > 4000000000ll) __trap(); }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t cols) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory"); }
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32 (bits 4-5 = 1), A and B tf32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at 17, M >> 4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { float e = __expf(-2.0f * fabsf(x)); float r = __fdividef(1.0f - e, 1.0f + e); return copysignf(r, x); }

// byte offset of element (row, k) inside an operand tile of `rows` rows (hi part; lo part follows at +2*rows*16)
__device__ __host__ __forceinline__ uint32_t op_off(int k, int row, int rows) { return (uint32_t)((k >> 3) * (4 * rows * 16) + ((k >> 2) & 1) * (rows * 16) + row * 16 + (k & 3) * 4); }
// nearest tf32 number (10 explicit mantissa bits): x = hi + lo with |lo| <= 2^-12 |x|, both exact in fp32
__device__ __host__ __forceinline__ float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
    uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & 0xFFFFE000u; float h; memcpy(&h, &b, 4); return h;
#endif
}
// split 4 consecutive k of one row into tf32-exact hi and the remainder lo, one float4 store each
__device__ __forceinline__ void store_hilo(unsigned char* tile, uint32_t off, uint32_t lo_off, const float (&v)[4]) {
    float4 hi, lo;
    hi.x = tf32_hi(v[0]); lo.x = v[0] - hi.x;
    hi.y = tf32_hi(v[1]); lo.y = v[1] - hi.y;
    hi.z = tf32_hi(v[2]); lo.z = v[2] - hi.z;
    hi.w = tf32_hi(v[3]); lo.w = v[3] - hi.w;
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + off + lo_off) = lo;
}

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tc
}  // namespace irrl
