// sm100_ptx.cuh -- thin inline-PTX layer for Blackwell (sm_100a): mbarrier, TMA (cp.async.bulk[.tensor]),
// tcgen05 (alloc / mma / commit / ld / st / fences) and the UMMA shared-memory + instruction descriptors.
// Nothing here is NRC-specific. Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptors"; the field
// positions were cross-checked against CUTLASS' cute/arch/mma_sm100_desc.hpp (read as documentation, no code taken).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
	uint32_t pred;
	asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
	return pred != 0;
}

// ---------------------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per warp: every lane orders its own tcgen05 traffic, the warp converges, lane 0 signals.
// (128 per-thread arrivals on one mbarrier serialise in the LSU; 4 per tile do not.)
__device__ __forceinline__ void warp_arrive_after_tcgen05(uint64_t *bar) {
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncwarp();
	if ((threadIdx.x & 31u) == 0)
		mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or `hint_ns` elapses (a hint, system-dependent).
// (measured on the inference kernel: no hint 67.4 us, any hint from 0 to 20 us 68.4 us, test_wait polling 73 us)
#ifndef SM100_MBAR_HINT_NS
#define SM100_MBAR_NO_HINT
#define SM100_MBAR_HINT_NS 0u
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity, uint32_t hint_ns = SM100_MBAR_HINT_NS) {
	uint32_t ok;
#if defined(SM100_MBAR_TEST_WAIT) // development switch: non-blocking poll instead of the hardware suspend
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok)
	             : "r"(smem_u32(bar)), "r"(parity)
	             : "memory");
#elif defined(SM100_MBAR_NO_HINT)
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok)
	             : "r"(smem_u32(bar)), "r"(parity)
	             : "memory");
#else
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok)
	             : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
	             : "memory");
#endif
	return ok != 0;
}
// Bounded wait: a protocol bug traps (reported by the C-ABI as a launch failure) instead of hanging the GPU.
#ifndef SM100_MBAR_SPIN_LIMIT
#define SM100_MBAR_SPIN_LIMIT (1u << 22)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	if (mbar_try_wait(bar, parity))
		return;
#pragma unroll 1 // never unroll: an unrolled spin loop bloats the kernel far past the instruction cache
	for (uint32_t i = 0; i < SM100_MBAR_SPIN_LIMIT; ++i)
		if (mbar_try_wait(bar, parity))
			return;
	__trap();
}

// A 32-bit counter in shared memory handed from one thread to pollers of the same CTA (buffer hand-backs that a parity wait
// could miss): release store / acquire load at CTA scope, i.e. everything the storing thread did before is visible to a
// poller that has seen the new value - and the pair is a synchronising access, not a data race, for the memory model and
// for compute-sanitizer's racecheck.
__device__ __forceinline__ void st_release_cta(volatile uint32_t *p, uint32_t v) {
	asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32((const void *)p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta(const volatile uint32_t *p) {
	uint32_t v;
	asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void *)p)) : "memory");
	return v;
}

// ---------------------------------------------------------------------------------------------------------- proxies
// generic-proxy writes to shared memory -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------- TMA
// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16 B aligned), completion on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
// 2-D tiled tensor copy global -> shared through a CUtensorMap (c0 = innermost coordinate).
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *tmap, int32_t c0, int32_t c1, uint64_t *bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
	                 "r"(smem_u32(smem_dst)),
	             "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
	asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---------------------------------------------------------------------------------------------------------- TMEM
// Tensor-memory address: bits [31:16] lane (data path), bits [15:0] column (32-bit cells).
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// whole warp; ncols power of two >= 32; result written to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
	             : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// all previously issued tcgen05.mma of this thread -> one arrival on `bar` when they complete
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues.
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	             "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	             : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
	             "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
	             : "memory");
}

// Same two forms with the shared-memory descriptors given as {low word, high word}. For the 128-byte-swizzled tiles used
// here only the 14-bit start-address field (low word) ever changes, so the issuing warp advances a descriptor with one
// 32-bit uniform add instead of 64-bit carry arithmetic in vector registers followed by R2UR moves.
constexpr uint32_t kSmemDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29); // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr) { return (smem_addr >> 4) & 0x3FFFu; } // LBO unused (0)
__device__ __forceinline__ void mma_ss_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
	             "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
	             "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
	             : "memory");
}
__device__ __forceinline__ void mma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
	             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
	             "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
	             : "memory");
}

// 32 data-path lanes x 32-bit, N consecutive columns: thread i of the warp <-> lane (taddr.lane + i).
// A warp may only touch the 32 lanes of its own sub-partition (warp_id % 4).
#define SM100_R4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
#define SM100_R16(v, o) SM100_R4(v, o), SM100_R4(v, o + 4), SM100_R4(v, o + 8), SM100_R4(v, o + 12)
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t (&v)[4]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : SM100_R4(v, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t *v) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	             : SM100_R16(v, 0)
	             : "r"(taddr)
	             : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t *v) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	             : SM100_R16(v, 0), SM100_R16(v, 16)
	             : "r"(taddr)
	             : "memory");
}
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t *v) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
	             "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
	             "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
	             : SM100_R16(v, 0), SM100_R16(v, 16), SM100_R16(v, 32), SM100_R16(v, 48)
	             : "r"(taddr)
	             : "memory");
}
// 64 consecutive columns holding 16-bit data (e.g. fp16 accumulators, one per 32-bit cell) -> 32 registers of packed pairs
__device__ __forceinline__ void tmem_ld_x32_pack16(uint32_t taddr, uint32_t *v) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	             : SM100_R16(v, 0), SM100_R16(v, 16)
	             : "r"(taddr)
	             : "memory");
}
#define SM100_I4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])
#define SM100_I16(v, o) SM100_I4(v, o), SM100_I4(v, o + 4), SM100_I4(v, o + 8), SM100_I4(v, o + 12)
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t *v) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), SM100_I4(v, 0),
	             SM100_I4(v, 4)
	             : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t *v) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
	             "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
	             SM100_I16(v, 0)
	             : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t *v) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
	             "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
	             "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
	             SM100_I16(v, 0), SM100_I16(v, 16)
	             : "memory");
}

// ---------------------------------------------------------------------------------------------------------- descriptors
// Instruction descriptor for kind::f16 with fp16 A/B and fp32 D.
//   [4,6) D format (1 = f32) | [7,10) A format (0 = f16) | [10,13) B format (0 = f16) | [15] A major (1 = MN)
//   [16] B major (1 = MN) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16_f32(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
	return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// same with fp16 D (accumulators rounded to fp16 after every K=16 instruction, one per 32-bit TMEM cell)
__host__ __device__ constexpr uint32_t make_idesc_f16_f16(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
	return make_idesc_f16_f32(M, N, a_mn_major, b_mn_major) & ~(3u << 4);
}
// relu on a packed pair of fp16 (HFMA2.RELU: fma pipe)
__device__ __forceinline__ uint32_t relu_f16x2(uint32_t x) {
	uint32_t r;
	asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0x3c003c00u), "r"(0x80008000u));
	return r;
}

// Shared-memory matrix descriptor, 128-byte swizzle. Every operand tile in this project is an array of 128-byte rows
// (64 fp16) starting on a 1024 B boundary, 16-byte chunk c of row r stored at chunk position c ^ (r & 7):
//   * K-major use  (K runs along the row):  8-row groups are SBO = 1024 B apart; a K=16 step advances the start by 32 B.
//   * MN-major use (K runs across rows):    8-row (= 8 k) groups are SBO = 1024 B apart; a K=16 step advances 2048 B;
//                                           LBO = distance between 64-element MN blocks (one block here -> unused).
//   [0,14) addr >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
	       ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

// byte offset of fp16 element (row, col) inside a 128B-swizzled tile of 128-byte rows
__host__ __device__ constexpr uint32_t sw128_offset(uint32_t row, uint32_t col) {
	return row * 128u + ((((col >> 3) ^ row) & 7u) << 4) + ((col & 7u) << 1);
}

// relu + fp32->fp16 (RNE) + pack in one instruction: low half <- lo, high half <- hi
__device__ __forceinline__ uint32_t cvt_relu_pack_f16x2(float lo, float hi) {
	uint32_t r;
	asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
	return r;
}
__device__ __forceinline__ uint32_t cvt_pack_f16x2(float lo, float hi) {
	uint32_t r;
	asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
	return r;
}

} // namespace sm100
