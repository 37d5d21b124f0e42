/*
 * crb_kernels.cuh -- the sm_100a tiled kernels of libclownresampler_b200.so (device code, templates).
 * Included by crb_device.cu (CUDA-runtime glue, direct kernel) and by crb_inst.cu, which instantiates one
 * kernel kind x channel range per translation unit so that the kinds build in parallel.
 *
 * The hot path replaced here is the reference's per-frame FIR (H:986-1035) inside its frame
 * loop (H:1058-1092), H = /root/reference/clownresampler.h.
 *
 * Kernel design (DESIGN.md has the long form):
 *  - Output frames are independent: frame n of a job sits at q(n) = q0 + n * increment (16.16,
 *    64-bit), q0 already including the radius delta, so its window starts at input frame
 *    ceil(q / 65536) and its phase is e = ceil(q/65536) * 65536 - q  (closed form of H:1076-1078,
 *    H:993-1001).
 *  - A CTA is persistent and walks tiles of `tile_out` consecutive output frames.  The input
 *    window of a tile (its frames plus the kernel-radius halo) is staged HBM -> shared memory
 *    with one 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx), double buffered so the
 *    copy of tile i+1 overlaps the arithmetic of tile i.  The per-phase tap table of the plan
 *    (crb_plan.c) lives in shared memory for the whole kernel.
 *  - Per tap and channel the reference adds trunc(s * k / 65536) (C division, toward zero,
 *    H:1020).  The kernel does that in ONE integer instruction: with S = s << 16 the 64-bit
 *    product S * k is p * 65536, and  hi32(S * k + (acc : bias))  = acc + floor((p*65536 + bias) / 2^32)
 *    equals acc + trunc(p / 65536) when bias = 0 for p >= 0 and 0xFFFFFFFF for p < 0 (IMAD.HI with
 *    a 64-bit addend whose low word carries the rounding bias).  The plan stores |k| and keeps
 *    positive- and negative-weight columns in separate chains, so sign(p) = sign(s) and the bias
 *    is just the sample's sign mask: one PRMT.  Bit-exact, no 64-bit accumulators needed
 *    (ranges proven per plan on the host).
 *  - Normalisation multiplies by the per-phase reciprocal 0x80000000 / sum(k) precomputed on the
 *    host with the reference's own integer division (H:1025) and truncates / 32768 (H:1033).
 */
#ifndef CRB_KERNELS_CUH
#define CRB_KERNELS_CUH
#include <cuda_runtime.h>

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "crb_internal.h"

#define CRB_INLINE_JOBS 8
#define CRB_STAGES CRB_RING_STAGES
#ifndef CRB_WAIT_HINT_NS
#define CRB_WAIT_HINT_NS 1000000u
#endif

struct crb_kparams {
	crb_geometry geo;
	const int32_t *rows;
	const int32_t *table;
	const crb_device_job *jobs;
	uint32_t n_jobs;
	uint32_t out_format;   /* 0 s32, 1 s16 clamped, 2 s32 raw accumulators + reciprocal */
	uint64_t total_tiles;
	crb_device_job inline_jobs[CRB_INLINE_JOBS];
#ifdef CRB_DEBUG_TIMING
	unsigned long long *dbg;
#endif
};

#ifdef CRB_DEBUG_TIMING
/* debug build only: [0] consumer-warp cycles waiting for a tile, [1] consumer-warp tiles, [2] producer cycles waiting for a free
   stage, [3] producer tiles, [4] consumer-warp cycles inside tiles, [5] producer cycles from stage-free to copies issued */
#endif

struct crb_tile_info {
	uint32_t t0;            /* (q - (ws0 - 1) * 65536) + 65535 for the tile's first frame: t >> 16 is 1 at input frame ws0 */
	uint32_t n_frames;      /* output frames of the tile, per stream */
	uint32_t increment;     /* 16.16 step of the tile's job */
	uint32_t n_streams;     /* 1, or 2 / 4 lockstep streams (unstretched kernel) */
	uint32_t win[CRB_MAX_LOCKSTEP];      /* per stream: shared-window address of input frame ws0 minus one frame, so that
	                                        win + (t >> 16) * frame_bytes is the first sample a frame reads */
	unsigned char *out[CRB_MAX_LOCKSTEP];   /* per stream: where the tile's first output frame goes */
};

/* ------------------------------------------------------------------------------------------
 * small PTX wrappers: mbarrier + 1-D TMA bulk copy (cp.async.bulk), sm_90+ forms valid on sm_100a
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	/* try_wait with a long suspend-time hint: the waiting warp sleeps in hardware until the phase completes instead of
	   polling (measured: without the hint the polls of waiting warps took one issue slot in eight from the working ones; an
	   explicit nanosleep between the polls on top of it changes nothing) */
#if CRB_WAIT_HINT_NS > 0
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"CRB_WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
		"@p bra CRB_DONE_%=;\n"
		"bra CRB_WAIT_%=;\n"
		"CRB_DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(CRB_WAIT_HINT_NS)
		: "memory");
#else
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"CRB_WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra CRB_DONE_%=;\n"
		"bra CRB_WAIT_%=;\n"
		"CRB_DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* ------------------------------------------------------------------------------------------
 * the exact multiply-accumulate: acc + trunc_toward_zero(s * k / 65536), k >= 0
 *   a * b must equal s * k * 65536 exactly:  (a, b) = (s << 16, k)   "big" columns, k up to 65536
 *                                            (a, b) = (s, k << 16)   "small" columns, k < 32768
 *   bias: any word whose top 16 bits are the sign of s -- the sign-extended sample itself.
 * hi32(a * b + (acc : bias)) = acc + floor((p * 65536 + bias) / 2^32), p = s * k:
 *   p >= 0: bias <= 0xFFFF never carries          -> acc + floor(p / 65536)
 *   p <  0: bias >= 0xFFFF0000 carries iff p is not a multiple of 65536 -> acc + ceil(p / 65536)
 * Compiles to one IMAD.HI Rd, a, b, (acc:bias).
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int mac_trunc(int acc, int a, int b, uint32_t bias)
{
	long long addend = (long long)(((unsigned long long)(uint32_t)acc << 32) | bias);
	/* Keep (acc : bias) opaque: otherwise ptxas re-associates the accumulator out of the 64-bit addend
	   (hi32(a*b + (0 : bias)) + acc), which costs a zeroing move and an add per MAC (measured 3-13 % slower). */
	asm("" : "+l"(addend));
	return (int)(((long long)a * (long long)b + addend) >> 32);
}

/* The same exact multiply-accumulate in three full-rate instructions instead of one quarter-rate IMAD.HI (any multiply with a
   64-bit product issues at about one per 4.4 clocks per scheduler on sm_100a and blocks the issue port meanwhile, measured:
   tools/microbench/overlap.cu): the product p = m * k fits 32 bits (|m| <= 32768, 0 <= k <= 65536), so
     t = m * k + (m < 0 ? 0xFFFF : 0)    one IMAD (the bias is one PRMT: the sign of byte 1 replicated into bytes 0 and 1)
     acc + (t >> 16)                     one LEA.HI.SX32
   floor((p + 65535) / 65536) = ceil(p / 65536) for p <= 0 and floor(p / 65536) for p >= 0: truncation toward zero (H:1020).
   `m` is the sign-extended sample, `k` the unshifted magnitude of the weight. */
__device__ __forceinline__ int mac_t16(int acc, int m, int k)
{
	uint32_t bias;
	asm("prmt.b32 %0, %1, %2, 0x4499;" : "=r"(bias) : "r"(m), "r"(0));
	const int t = m * k + (int)bias;
	return acc + (t >> 16);
}

/* ... and in TWO instructions when the running sum of a chain is known to fit 16 bits: the accumulator lives in the upper
   half of `acc` (the lower half is garbage), one PRMT builds the addend { upper half of acc : 0xFFFF when m < 0 else 0 } and
   one IMAD adds the product: acc' = m * k + addend = 65536 * (sum + trunc(m * k / 65536)) + remainder.  The caller reads the
   sum with one arithmetic shift when the chain is complete. */
__device__ __forceinline__ int mac_hi16(int acc, int m, int k)
{
	uint32_t addend;
	asm("prmt.b32 %0, %1, %2, 0x7699;" : "=r"(addend) : "r"(m), "r"(acc));
	return m * k + (int)addend;
}

/* out = trunc(acc * recip / 32768) (H:1033) in the same three-instruction form, for plans whose reciprocals sit close to
   32768 (proven per plan: |acc * 2 * (recip - 32768)| + 65535 < 2^31).  acc * recip / 32768 = acc + acc * rd2 / 65536 with
   rd2 = 2 * (recip - 32768); the sum has the sign of acc, so truncation toward zero is acc + floor(x) for acc >= 0 and
   acc + ceil(x) for acc < 0: acc + ((acc * rd2 + (acc < 0 ? 0xFFFF : 0)) >> 16). */
__device__ __forceinline__ int normalise_t16(int acc, int rd2)
{
	uint32_t bias;
	asm("prmt.b32 %0, %1, %2, 0x44BB;" : "=r"(bias) : "r"(acc), "r"(0));
	const int t = acc * rd2 + (int)bias;
	return acc + (t >> 16);
}

/* PTX prmt in default mode: selector nibble bit 3 replicates the sign bit of the selected byte
   (the CUDA intrinsic __byte_perm masks that bit off, so it has to be inline PTX). */
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}

/* shared-memory loads by 32-bit shared-window address (keeps the address arithmetic 32-bit) */
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ int lds_s16(uint32_t addr) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }   /* sign-extending load */
__device__ __forceinline__ uint2 lds64(uint32_t addr) { uint2 v; asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t addr) { uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v; }

/* One packed word = two s16 samples (lo = even channel, hi = odd channel). */
/* SIGNED: the column holds the signed weight (its sign differs between phase rows); the product then has the sign of
   s ^ k, so the bias is the sample with all bits flipped when k < 0 (s > 0: 0xFFFF8000..0xFFFFFFFE, s < 0: 0..0x7FFF,
   s == 0 or k == 0: a zero product, which no bias can carry).  `ks` = k >> 31, computed once per tap. */
template <bool BIG, bool SIGNED>
__device__ __forceinline__ void tap_word(int &acc_lo, int &acc_hi, uint32_t w, int k, uint32_t ks)
{
	const int m_lo = (int)prmt(w, 0, 0x9910);   /* sign-extended low half: multiplicand of small columns, bias of all */
	const int m_hi = (int)w >> 16;
	const uint32_t b_lo = SIGNED ? (uint32_t)m_lo ^ ks : (uint32_t)m_lo, b_hi = SIGNED ? (uint32_t)m_hi ^ ks : (uint32_t)m_hi;
	if (BIG) {
		acc_lo = mac_trunc(acc_lo, (int)prmt(w, 0, 0x1044), k, b_lo);   /* w << 16, kept off the multiplier pipe */
		acc_hi = mac_trunc(acc_hi, (int)(w & 0xFFFF0000u), k, b_hi);
	} else {
		acc_lo = mac_trunc(acc_lo, m_lo, k, b_lo);
		acc_hi = mac_trunc(acc_hi, m_hi, k, b_hi);
	}
}

template <bool BIG, bool SIGNED>
__device__ __forceinline__ int tap_scalar(int acc, int m, int k, uint32_t ks)
{
	return mac_trunc(acc, BIG ? m << 16 : m, k, SIGNED ? (uint32_t)m ^ ks : (uint32_t)m);
}

/* SPLIT (stereo): two sign-extending 16-bit loads instead of one 32-bit load and two unpack operations.  In the
   unstretched kernel the ALU pipe is fuller than the shared-memory pipe (measured 4.6 % faster on 44.1 -> 48 kHz);
   the general kernel is bound by its load count and keeps the packed load (measured 40 % slower with SPLIT). */
template <int C, bool BIG, bool SPLIT = false, bool SIGNED = false>
__device__ __forceinline__ void tap(int (&acc)[16], uint32_t frame, int k, int channels)
{
	const uint32_t ks = SIGNED ? (uint32_t)(k >> 31) : 0u;
	if (C == 1) {
		acc[0] = tap_scalar<BIG, SIGNED>(acc[0], lds_s16(frame), k, ks);
	} else if (C == 2 && SPLIT) {
		const int m0 = lds_s16(frame), m1 = lds_s16(frame + 2);
		/* m << 16 as a byte permute: keeps the shift on the ALU pipe (1 % faster than leaving the choice to ptxas) */
		acc[0] = BIG ? mac_trunc(acc[0], (int)prmt((uint32_t)m0, 0, 0x1044), k, (uint32_t)m0) : mac_trunc(acc[0], m0, k, (uint32_t)m0);
		acc[1] = BIG ? mac_trunc(acc[1], (int)prmt((uint32_t)m1, 0, 0x1044), k, (uint32_t)m1) : mac_trunc(acc[1], m1, k, (uint32_t)m1);
	} else if (C == 2) {
		tap_word<BIG, SIGNED>(acc[0], acc[1], lds32(frame), k, ks);
	} else if (C == 4) {
		const uint2 v = lds64(frame);
		tap_word<BIG, SIGNED>(acc[0], acc[1], v.x, k, ks);
		tap_word<BIG, SIGNED>(acc[2], acc[3], v.y, k, ks);
	} else if (C == 8) {
		const uint4 v = lds128(frame);
		tap_word<BIG, SIGNED>(acc[0], acc[1], v.x, k, ks);
		tap_word<BIG, SIGNED>(acc[2], acc[3], v.y, k, ks);
		tap_word<BIG, SIGNED>(acc[4], acc[5], v.z, k, ks);
		tap_word<BIG, SIGNED>(acc[6], acc[7], v.w, k, ks);
	} else if (C == 6) {
		/* 12-byte frames are 4-byte aligned: three packed loads */
		tap_word<BIG, SIGNED>(acc[0], acc[1], lds32(frame), k, ks);
		tap_word<BIG, SIGNED>(acc[2], acc[3], lds32(frame + 4), k, ks);
		tap_word<BIG, SIGNED>(acc[4], acc[5], lds32(frame + 8), k, ks);
	} else {
		/* odd channel counts (compile-time C = 3, 5, 7) and C == 0 (9..16 channels, count at run time): scalar 16-bit loads */
#pragma unroll
		for (int c = 0; c < 16; ++c)
			if (c < channels)
				acc[c] = tap_scalar<BIG, SIGNED>(acc[c], lds_s16(frame + 2 * c), k, ks);
	}
}

/* out = trunc(acc * recip / 32768), H:1033, as one IMAD.HI when the plan allows it.  With rd = recip - 32768:
     mode 3: row word = rd << 17:  hi32(acc        * word + (acc : acc))       (|rd| < 16384, |acc| < 2^17)
     mode 2: row word = rd << 17:  hi32(acc        * word + (acc : acc >> 31)) (|rd| < 16384)
     mode 1: row word = rd << 16:  hi32((acc << 1) * word + (acc : acc >> 31)) (recip < 65536, |acc| < 2^30)
   all equal floor((acc * recip * 2^17 + bias) / 2^32) with a bias whose top 15 bits are the sign of acc and
   whose value is below 2^17 for acc >= 0: truncation toward zero by the argument of mac_trunc.
     mode 0: row word = recip, plain 64-bit arithmetic. */
__device__ __forceinline__ int normalise(int acc, int row_word, uint32_t mode)
{
	/* acc usually is the high half of the previous 64-bit multiply-add: keep it a 32-bit value, or the compiler
	   multiplies the un-truncated 64-bit intermediate instead (a 64 x 64 product in four pieces) */
	asm("" : "+r"(acc));
	if (mode == 3)
		return mac_trunc(acc, acc, row_word, (uint32_t)acc);
	if (mode == 2)
		return mac_trunc(acc, acc, row_word, (uint32_t)(acc >> 31));
	if (mode == 1)
		return mac_trunc(acc, acc << 1, row_word, (uint32_t)(acc >> 31));
	const long long q = (long long)acc * (long long)row_word;
	return (int)((q + ((q >> 63) & 32767)) >> 15);
}

__device__ __forceinline__ int recip_of_row_word(int row_word, uint32_t mode)
{
	return mode >= 2 ? (row_word >> 17) + 32768 : mode == 1 ? (row_word >> 16) + 32768 : row_word;
}

__device__ __forceinline__ int clamp_s16(int v)
{
	return max(-0x7FFF, min(0x7FFF, v));   /* examples/low-level.c:74-77 */
}

/* two channels: saturating pack to s16x2 (I2IP.S16.S32.SAT clamps to [-32768, 32767]) then a packed max
   with -32767 (VIMNMX.S16x2) gives the reference callbacks' [-0x7FFF, 0x7FFF] clamp in two instructions */
__device__ __forceinline__ uint32_t clamp_pack2(int lo, int hi)
{
	uint32_t r;
	asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
	return __vmaxs2(r, 0x80018001u);
}

template <int C, int FMT>
__device__ __forceinline__ void store_frame(unsigned char *out, const int (&v)[16], int channels, int recip)
{
	if (FMT == 2) {
		/* diagnostic format: un-normalised accumulators followed by the phase reciprocal */
		int *o = (int *)out;
#pragma unroll
		for (int c = 0; c < 16; ++c) if (c < channels) o[c] = v[c];
		o[channels] = recip;
	} else if (FMT == 0) {
		int *o = (int *)out;
		if (C == 2) { *(int2 *)o = make_int2(v[0], v[1]); }
		else if (C == 4) { *(int4 *)o = make_int4(v[0], v[1], v[2], v[3]); }
		else if (C == 8) { ((int4 *)o)[0] = make_int4(v[0], v[1], v[2], v[3]); ((int4 *)o)[1] = make_int4(v[4], v[5], v[6], v[7]); }
		else {
#pragma unroll
			for (int c = 0; c < 16; ++c) if (c < channels) o[c] = v[c];
		}
	} else {
		int16_t *o = (int16_t *)out;
		if (C == 2) {
			*(uint32_t *)o = clamp_pack2(v[0], v[1]);
		} else if (C == 4) {
			*(uint2 *)o = make_uint2(clamp_pack2(v[0], v[1]), clamp_pack2(v[2], v[3]));
		} else if (C == 8) {
			*(uint4 *)o = make_uint4(clamp_pack2(v[0], v[1]), clamp_pack2(v[2], v[3]), clamp_pack2(v[4], v[5]), clamp_pack2(v[6], v[7]));
		} else {
#pragma unroll
			for (int c = 0; c < 16; ++c) if (c < channels) o[c] = (int16_t)clamp_s16(v[c]);
		}
	}
}

__device__ __forceinline__ const crb_device_job *job_table(const crb_kparams &p)
{
	return p.jobs ? p.jobs : p.inline_jobs;
}

/* last job whose tile_base <= tile */
__device__ __forceinline__ uint32_t find_job(const crb_kparams &p, uint64_t tile)
{
	const crb_device_job *jobs = job_table(p);
	uint32_t lo = 0, hi = p.n_jobs;
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (jobs[mid].tile_base <= tile) lo = mid; else hi = mid;
	}
	return lo;
}

__device__ __forceinline__ uint32_t out_frame_bytes(const crb_kparams &p)
{
	return p.out_format == 1 ? p.geo.channels * 2u : (p.geo.channels + (p.out_format == 2)) * 4u;
}

/* Producer warp: describe tile `tile` of job `job`, and start the bulk copies of its input windows into `stage`.  Lane s
   (s < streams of the job) looks after lockstep stream s, lane 0 also writes the tile header and arms the barrier.  A
   window starts at the 16-byte boundary at or below the first frame the tile reads; the lead that rounding leaves in
   front of that frame goes into the stream's window address.  All address arithmetic happens BEFORE the wait for the
   stage to be released, so that the copies start the moment the consumers let go of it. */
__device__ __forceinline__ void produce_tile(const crb_kparams &p, const crb_device_job &job, uint64_t tile, unsigned char *stage, crb_tile_info *info, uint64_t *bar,
	bool wait_empty, uint64_t *empty_bar, uint32_t empty_parity, uint32_t lane
#ifdef CRB_DEBUG_TIMING
	, unsigned long long (&dbg_acc)[3]
#endif
	)
{
	const crb_geometry &g = p.geo;
	const uint32_t n_streams = 1u + job.n_more;
	const uint32_t log_streams = n_streams >> 1;                       /* 1, 2, 4 -> 0, 1, 2 */
	const uint32_t tile_out = g.tile_out >> log_streams;
	const uint32_t slot_bytes = n_streams == 1 ? 0u : g.lock_slot_bytes[log_streams];
	const uint64_t first = (tile - job.tile_base) * tile_out;
	const uint64_t left = job.n_out - first;
	const uint32_t n = left < tile_out ? (uint32_t)left : tile_out;
	const uint64_t inc = job.increment ? job.increment : g.increment;
	const uint64_t q = job.q0 + (job.first_out + first) * inc;
	const uint64_t ws0 = (q + 65535) >> 16;
	const uint64_t ws_last = (q + (uint64_t)(n - 1) * inc + 65535) >> 16;
	const uint32_t frame_bytes = 2 * g.channels;
	const bool mine = lane < n_streams;
	const uint32_t s = mine ? lane : 0u;
	unsigned char *slot = stage + s * slot_bytes;
	const int16_t *in = job.in;
	void *out = job.out;
	uint64_t in_frames = job.in_frames;
#pragma unroll
	for (uint32_t m = 1; m < CRB_MAX_LOCKSTEP; ++m)
		if (s == m) { in = job.in_more[m - 1]; out = job.out_more[m - 1]; in_frames = job.in_frames_more[m - 1]; }
	uint64_t end_frame = ws_last + g.taps_max;
	if (end_frame > in_frames) end_frame = in_frames;   /* columns past the buffer end are zero-weight */
	const uintptr_t a_first = (uintptr_t)in + ws0 * frame_bytes;
	const uintptr_t a_end = (uintptr_t)in + end_frame * frame_bytes;
	const uintptr_t a0 = a_first & ~(uintptr_t)15;
	/* The bulk copy moves whole 16-byte chunks.  Its start may round down into the chunk that holds the first frame
	   (same allocation); its end must not round up past the caller's buffer: when the last chunk is ragged
	   (buffer end not 16-byte aligned) the chunk is copied by this lane with 2-byte loads instead. */
	const uintptr_t buf_end = (uintptr_t)in + in_frames * frame_bytes;
	uintptr_t a1 = (a_end + 15) & ~(uintptr_t)15;
	const bool ragged = a1 > buf_end;
	if (ragged) {
		a1 = a_end & ~(uintptr_t)15;
		if (a1 < a0) a1 = a0;
	}
	const uint32_t bytes = mine ? (uint32_t)(a1 - a0) : 0u;
	uint32_t total_bytes = bytes;
#pragma unroll
	for (uint32_t o = 1; o < CRB_MAX_LOCKSTEP; o <<= 1) total_bytes += __shfl_xor_sync(0xFFFFFFFFu, total_bytes, o);

	/* the tile descriptor lives in a ring twice as deep as the stage ring: it is written while the consumers still work
	   on the stage's previous tile, so that after the wait only the barrier arming and the copies remain */
	if (mine) {
		info->win[s] = smem_u32(slot) + (uint32_t)(a_first - a0) - frame_bytes;
		info->out[s] = (unsigned char *)out + first * out_frame_bytes(p);
	}
	if (lane == 0) {
		info->t0 = (uint32_t)(q - ((ws0 - 1) << 16)) + 65535u;   /* t0 >> 16 == 1 at frame ws0 */
		info->n_frames = n;
		info->increment = (uint32_t)inc;
		info->n_streams = n_streams;
	}
	__syncwarp();
#ifdef CRB_DEBUG_TIMING
	const long long dbg_t0 = clock64();
#endif
	/* After the wait only the barrier arming and the copies remain, each lane on its own: no warp-wide synchronisation on
	   the path that decides how long the consumers wait for their next tile.  (A copy may complete its bytes on the
	   barrier before lane 0 has announced them: the transaction count just goes negative for a moment; the phase cannot
	   complete before lane 0's arrival.)  Only a ragged tail -- the last tile of a buffer whose end is not 16-byte
	   aligned -- takes the ordered path, because its 2-byte stores must be released by lane 0's arrival. */
	const bool any_ragged = __any_sync(0xFFFFFFFFu, mine && ragged);
	if (wait_empty && (mine || any_ragged)) mbar_wait(empty_bar, empty_parity);
#ifdef CRB_DEBUG_TIMING
	const long long dbg_t1 = clock64();
#endif
	if (any_ragged) {
		if (mine && ragged) {
			const uint16_t *from = (const uint16_t *)(a1 > a_first ? a1 : a_first);
			uint16_t *to = (uint16_t *)(slot + ((uintptr_t)from - a0));
			for (; (uintptr_t)from < a_end; ++from, ++to) *to = *from;
		}
		__syncwarp();
	}
	/* lane 0's arrive has release semantics; the descriptor writes of the other lanes were ordered before it by the
	   __syncwarp above the wait */
	if (lane == 0) mbar_arrive_expect_tx(bar, total_bytes);
	if (bytes) tma_bulk_g2s(slot, (const void *)a0, bytes, bar);
#ifdef CRB_DEBUG_TIMING
	const long long dbg_t2 = clock64();
	dbg_acc[0] += (unsigned long long)(dbg_t1 - dbg_t0); dbg_acc[1] += 1ull; dbg_acc[2] += (unsigned long long)(dbg_t2 - dbg_t1);
#endif
}

/* ------------------------------------------------------------------------------------------
 * the tiled kernel
 *   C    : compile-time channel count 1..8 (packed vector loads for 2, 4, 6, 8), or 0 = count at run time (9..16), scalar loads
 *   FMT  : 0 = s32 unclamped, 1 = s16 clamped, 2 = raw accumulators + reciprocal
 *   K    : 1 = unstretched 5-column kernel with compile-time signs + - + + - and packed 16-byte rows,
 *          6 / 8 / 10 / 12 = slightly stretched kernel unrolled over that many signed taps, 0 = general kernel
 *
 * CRB_NT_K(C, K == 1) / 32 consumer warps (20, 16 or 8) + 1 producer warp, CRB_STAGES-deep ring of input windows:
 *   producer lane : wait empty[s] -> describe tile, arm full[s] with the byte count, issue the TMA bulk copy
 *   consumer warp : wait full[s]  -> its frames of the tile -> arrive on empty[s]
 * No CTA-wide barrier in steady state.
 * ------------------------------------------------------------------------------------------ */
/* One output frame of the unstretched 5-column kernel.
   `t`     : tile-relative position word; t >> 16 = window start in stage frames (1-based), ~t & 0xFFFF = phase
   `stage` : shared address of the tile's input window minus one frame (plus the lead for odd frame sizes)
   `rows`  : shared address of the packed table, 16 bytes per phase row:
             { k2, k3, (k1 << 16) | k0, (k4 << 16) | (2 * (recip - 32768) & 0xFFFF) } */
__device__ __forceinline__ uint4 u5_row(uint32_t t, uint32_t rows)
{
	return lds128(rows + (~(t >> 2) & 0x3FF0u));
}

template <int C, bool SINGLE, bool SIGNED>
__device__ __forceinline__ void chain_tap(int (&a)[16], uint32_t frame, int k, int channels);

/* The arithmetic of one unstretched frame given its phase row `r`: five taps with the static signs + - + + -, in three
   16-bit chains (mac_hi16) for 1, 2 and odd channel counts (and the run-time count), which load every sample with its own
   sign-extending 16-bit load; 4, 6 and 8 channels keep packed vector loads + IMAD.HI. */
template <int C, int FMT>
__device__ __forceinline__ void frame_u5_row(const uint4 r, uint32_t win, unsigned char *outp, int channels)
{
	const uint32_t fb = 2u * channels;
	int accp[16], accn[16], outv[16];
#pragma unroll
	for (int c = 0; c < 16; ++c) accp[c] = accn[c] = 0;
	if (C == 4 || C == 6 || C == 8) {
#ifndef CRB_U5_PACKED_CHAINS
		/* measured: the 16-bit chains on packed words (-DCRB_U5_PACKED_CHAINS) are +3 % on 4 channels, -3 % on 6, equal on 8 */
		tap<C, false>(accp, win, (int)(r.z << 16), channels);
		tap<C, false>(accn, win + fb, (int)(r.z & 0xFFFF0000u), channels);
		tap<C, true>(accp, win + 2 * fb, (int)r.x, channels);
		tap<C, true>(accp, win + 3 * fb, (int)r.y, channels);
		tap<C, false>(accn, win + 4 * fb, (int)(r.w & 0xFFFF0000u), channels);
#else
		/* the same three chains as below, on packed words */
		int x[16], y[16], z[16];
#pragma unroll
		for (int c = 0; c < 16; ++c) x[c] = y[c] = z[c] = 0;
		chain_tap<C, false, false>(x, win + 3 * fb, (int)r.y, channels);
		chain_tap<C, false, false>(x, win, (int)(r.z & 0xFFFFu), channels);
		chain_tap<C, false, false>(y, win + 2 * fb, (int)r.x, channels);
		chain_tap<C, false, false>(z, win + fb, (int)(r.z >> 16), channels);
		chain_tap<C, false, false>(z, win + 4 * fb, (int)(r.w >> 16), channels);
#pragma unroll
		for (int c = 0; c < 16; ++c)
			if (c < channels) { accp[c] = (x[c] >> 16) + (y[c] >> 16); accn[c] = z[c] >> 16; }
#endif
	} else {
		/* Three accumulator chains per channel, each kept in the UPPER half of a register (mac_hi16): {tap 3, tap 0}, {tap 2} and
		   the negative taps {1, 4}.  The plan proves per phase row that the weights of a chain sum to at most 65536, so a chain
		   stays inside 16 bits whatever the samples are; the three are combined in 32 bits. */
		const int k0 = (int)(r.z & 0xFFFFu), k1 = (int)(r.z >> 16), k4 = (int)(r.w >> 16);
#pragma unroll
		for (int c = 0; c < 16; ++c)
			if (c < channels) {
				int x = mac_hi16(0, lds_s16(win + 3 * fb + 2 * c), (int)r.y);
				x = mac_hi16(x, lds_s16(win + 2 * c), k0);
				const int y = mac_hi16(0, lds_s16(win + 2 * fb + 2 * c), (int)r.x);
				int z = mac_hi16(0, lds_s16(win + fb + 2 * c), k1);
				z = mac_hi16(z, lds_s16(win + 4 * fb + 2 * c), k4);
				accp[c] = (x >> 16) + (y >> 16);
				accn[c] = z >> 16;
			}
	}
	const int rd2 = (int)prmt(r.w, 0, 0x9910);       /* 2 * (recip - 32768), sign-extended from the low half */
#pragma unroll
	for (int c = 0; c < 16; ++c)
		if (c < channels) outv[c] = FMT == 2 ? accp[c] - accn[c] : normalise_t16(accp[c] - accn[c], rd2);
	store_frame<C, FMT>(outp, outv, channels, (rd2 >> 1) + 32768);
}

/* Mono, unstretched: two ADJACENT output frames from one six-sample window.  Up-sampling steps by at most one input
   frame per output frame, so frame B's window starts at frame A's or one sample later: six loads serve both frames
   (three per frame instead of five), five selects align B.  `ra`, `rb`: the phase rows of the two frames. */
template <int FMT>
__device__ __forceinline__ void frame_u5_mono_pair(const uint4 ra, const uint4 rb, bool shifted, uint32_t win, unsigned char *outp)
{
	int s[6];
#pragma unroll
	for (int j = 0; j < 6; ++j) s[j] = lds_s16(win + 2 * j);
	int pair[2];
#pragma unroll
	for (int f = 0; f < 2; ++f) {
		const uint4 r = f ? rb : ra;
		int x[5];
#pragma unroll
		for (int j = 0; j < 5; ++j) x[j] = (f && shifted) ? s[j + 1] : s[j];
		int cx = mac_hi16(0, x[3], (int)r.y);
		cx = mac_hi16(cx, x[0], (int)(r.z & 0xFFFFu));
		const int cy = mac_hi16(0, x[2], (int)r.x);
		int cz = mac_hi16(0, x[1], (int)(r.z >> 16));
		cz = mac_hi16(cz, x[4], (int)(r.w >> 16));
		const int accp = (cx >> 16) + (cy >> 16), accn = cz >> 16;
		const int rd2 = (int)prmt(r.w, 0, 0x9910);
		pair[f] = FMT == 2 ? accp - accn : normalise_t16(accp - accn, rd2);
		if (FMT != 1) {
			int outv[16];
			outv[0] = pair[f];
			store_frame<1, FMT>(outp + f * (FMT == 2 ? 8u : 4u), outv, 1, (rd2 >> 1) + 32768);
		}
	}
	/* s16: the two frames are neighbours in memory: one saturating pack, one packed clamp and one 32-bit store for both
	   (the caller takes this path only for 4-byte aligned pairs) */
	if (FMT == 1) *(uint32_t *)outp = clamp_pack2(pair[0], pair[1]);
}

/* A full tile of the unstretched kernel: FULL_TILE / G frames of each of G lockstep streams (G = 1, 2, 4), 16 frame
   computations per thread, fully unrolled, stores at immediate offsets.  Thread `tid` takes frames tid, tid + NT, ... of
   every stream; a frame's phase row is fetched once and serves that frame of all G streams. */
template <int C, int FMT, int G, uint32_t NT, uint32_t FULL_TILE>
__device__ __forceinline__ void u5_full_tile(const crb_tile_info &info, uint32_t tid, uint32_t rows, uint32_t fb_out, int channels)
{
	const uint32_t fb = 2u * channels;
	uint32_t win[G];
	unsigned char *out[G];
#pragma unroll
	for (int s = 0; s < G; ++s) { win[s] = info.win[s]; out[s] = info.out[s]; }
	const uint32_t t_step = NT * info.increment;
	if (C == 1) {
		/* mono: thread `tid` takes the frame pairs tid, tid + NT, ... (frames 2p and 2p + 1) */
		const uint32_t tp = info.t0 + 2u * tid * info.increment;
#pragma unroll
		for (int k = 0; k < (int)(FULL_TILE / NT / 2 / G); ++k) {
			const uint32_t ta = tp + 2u * k * t_step, tb = ta + info.increment;
			const uint4 ra = u5_row(ta, rows), rb = u5_row(tb, rows);
			const bool shifted = (tb >> 16) != (ta >> 16);
#pragma unroll
			for (int s = 0; s < G; ++s)
				frame_u5_mono_pair<FMT>(ra, rb, shifted, win[s] + (ta >> 16) * 2u, out[s] + ((size_t)2 * tid + (size_t)2 * k * NT) * fb_out);
		}
	} else {
		const uint32_t t = info.t0 + tid * info.increment;
#pragma unroll
		for (int k = 0; k < (int)(FULL_TILE / NT / G); ++k) {
			const uint32_t tk = t + k * t_step;
			const uint4 r = u5_row(tk, rows);
#pragma unroll
			for (int s = 0; s < G; ++s)
				frame_u5_row<C, FMT>(r, win[s] + (tk >> 16) * fb, out[s] + ((size_t)tid + (size_t)k * NT) * fb_out, channels);
		}
	}
}

/* Two columns of one group: weights {k0, k1} and frame byte offsets {o0, o1} arrive in two 64-bit loads. */
/* CONSTOFF: `ci` is the column INDEX and the offsets come from the kernel parameters (uniform across the warp: the
   compiler keeps them in uniform registers and folds them into the load address); otherwise `ci` is the shared
   address of the column's offset word (rotating plans: columns differ per lane). */
template <int C, bool BIG, bool SIGNED, bool CONSTOFF>
__device__ __forceinline__ void pair_taps(const crb_geometry &g, int (&acc)[16], uint32_t w, uint32_t ci, uint32_t win, int channels)
{
	const uint2 kk = lds64(w);
	uint32_t o0, o1;
	if (CONSTOFF) { o0 = g.col_off16[ci]; o1 = g.col_off16[ci + 1]; }
	else { const uint2 oo = lds64(ci); o0 = oo.x; o1 = oo.y; }
	tap<C, BIG, false, SIGNED>(acc, win + o0, (int)kk.x, channels);
	tap<C, BIG, false, SIGNED>(acc, win + o1, (int)kk.y, channels);
}

/* One column group (`count` columns, even): the pairs beyond a multiple of four run as straight-line code first
   (a compiler-generated remainder loop would run them one by one, without overlap), then four pairs per iteration. */
template <int C, bool BIG, bool SIGNED, bool CONSTOFF>
__device__ __forceinline__ void group_taps(const crb_geometry &g, int (&acc)[16], uint32_t w, uint32_t ci, uint32_t win, uint32_t count, int channels)
{
	constexpr uint32_t CS = CONSTOFF ? 2u : 8u;     /* step of `ci` per pair: two columns, or two offset words */
	const uint32_t pairs = count >> 1, rem = pairs & 3u;
	if (rem == 3) {
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w, ci, win, channels);
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w + 8, ci + CS, win, channels);
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w + 16, ci + 2 * CS, win, channels);
	} else if (rem == 2) {
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w, ci, win, channels);
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w + 8, ci + CS, win, channels);
	} else if (rem == 1) {
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w, ci, win, channels);
	}
	w += rem * 8;
	ci += rem * CS;
#pragma unroll 1
	for (uint32_t i = rem; i < pairs; i += 4, w += 32, ci += 4 * CS) {
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w, ci, win, channels);
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w + 8, ci + CS, win, channels);
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w + 16, ci + 2 * CS, win, channels);
		pair_taps<C, BIG, SIGNED, CONSTOFF>(g, acc, w + 24, ci + 3 * CS, win, channels);
	}
}

/* One output frame of the general kernel: phase row by the plan's formula, then the plan's six column groups
   (positive, negative, signed) x (small, big). */
template <int C, int FMT>
__device__ __forceinline__ void frame_runs(const crb_geometry &g, uint32_t t, uint32_t stage, uint32_t rows, unsigned char *outp, int channels, uint32_t lane_rot)
{
	const uint32_t fb = 2u * channels;
	const uint32_t e = ~t & 0xFFFFu;
	uint32_t r = (((e + g.delta) * g.step) >> 16) - g.ks0;
#pragma unroll
	for (uint32_t b = 0; b < CRB_MAX_BREAKS; ++b) r += (e >= g.breaks[b]);   /* unused thresholds are 0xFFFFFFFF */
	const uint32_t row = rows + r * g.row_words * 4;
	const uint32_t colinfo = rows + g.n_rows * g.row_words * 4;
	const uint32_t win = stage + (t >> 16) * fb;
	int accp[16], accn[16], outv[16];
#pragma unroll
	for (int c = 0; c < 16; ++c) accp[c] = accn[c] = 0;
	/* lane_rot: byte offset of this lane's first pair inside a rotating group (the group is followed by a copy of
	   its first columns, so the loop runs straight through) */
#define CRB_GROUP(G, ACC, BIG, SIGNED, CONSTOFF) \
	if (g.groups[G][1]) { \
		const uint32_t o = g.groups[G][0] * 4 + (lane_rot & g.group_rot[G]); \
		group_taps<C, BIG, SIGNED, CONSTOFF>(g, ACC, row + o, CONSTOFF ? g.groups[G][0] : colinfo + o, win, g.groups[G][1], channels); \
	}
#define CRB_ALL_GROUPS(CONSTOFF) \
	CRB_GROUP(0, accp, false, false, CONSTOFF) \
	CRB_GROUP(1, accp, true, false, CONSTOFF) \
	CRB_GROUP(2, accn, false, false, CONSTOFF) \
	CRB_GROUP(3, accn, true, false, CONSTOFF) \
	CRB_GROUP(4, accp, false, true, CONSTOFF) \
	CRB_GROUP(5, accp, true, true, CONSTOFF)
	/* measured: offsets through the constant cache are 4-10 % faster with scalar sample loads (1, 3, 5, 7 channels)
	   and 2-8 % slower with packed loads, so only the odd instantiations carry that path */
	if ((C & 1) && g.const_offsets) { CRB_ALL_GROUPS(true) } else { CRB_ALL_GROUPS(false) }
#undef CRB_ALL_GROUPS
#undef CRB_GROUP
	const int recip_word = (int)lds32(row + g.n_cols * 4);
	/* one (warp-uniform) branch on the plan's normaliser form, not one per channel */
#define CRB_NORMALISE_ALL(MODE) \
	_Pragma("unroll") for (int c = 0; c < 16; ++c) if (c < channels) outv[c] = normalise(accp[c] - accn[c], recip_word, MODE);
	if (FMT == 2) {
#pragma unroll
		for (int c = 0; c < 16; ++c) if (c < channels) outv[c] = accp[c] - accn[c];
	} else if (g.norm_mode == 1) { CRB_NORMALISE_ALL(1)
	} else if (g.norm_mode == 3) { CRB_NORMALISE_ALL(3)
	} else if (g.norm_mode == 2) { CRB_NORMALISE_ALL(2)
	} else { CRB_NORMALISE_ALL(0) }
#undef CRB_NORMALISE_ALL
	store_frame<C, FMT>(outp, outv, channels, recip_of_row_word(recip_word, g.norm_mode));
}


/* ---- general kernel, chain form (crb_geometry.chain_mode): no 64-bit-product multiply anywhere in the tap loop ----
   A column's weight is the plain |k| (positive / negative class) or the signed k (signed class).  One tap of a CHAIN group
   costs, per sample, the extraction of the sample from its packed word, one PRMT and one IMAD (mac_hi16: the chain's running
   sum lives in the upper 16 bits, the plan proves it fits); a SINGLE column costs one more (mac_t16).  For the signed class the
   sign that picks the rounding direction is the sample's XOR the weight's: the PRMTs read it from the packed word flipped by
   ks = k >> 31. */
template <bool SINGLE, bool SIGNED>
__device__ __forceinline__ void chain_word(int &a_lo, int &a_hi, uint32_t w, int k, uint32_t ks)
{
	const int m_lo = (int)prmt(w, 0, 0x9910);   /* sign-extended low half */
	const int m_hi = (int)w >> 16;
	const uint32_t src = SIGNED ? w ^ ks : w;
	if (SINGLE) {
		const int t_lo = m_lo * k + (int)prmt(src, 0, 0x4499);     /* + 0xFFFF when the product is negative */
		const int t_hi = m_hi * k + (int)prmt(src, 0, 0x44BB);
		a_lo += t_lo >> 16;
		a_hi += t_hi >> 16;
	} else {
		a_lo = m_lo * k + (int)prmt(src, (uint32_t)a_lo, 0x7699);  /* { upper half of the chain : 0xFFFF when the product is negative } */
		a_hi = m_hi * k + (int)prmt(src, (uint32_t)a_hi, 0x76BB);
	}
}

template <bool SINGLE, bool SIGNED>
__device__ __forceinline__ int chain_scalar(int a, int m, int k, uint32_t ks)
{
	const uint32_t src = SIGNED ? (uint32_t)m ^ ks : (uint32_t)m;
	if (SINGLE) {
		const int t = m * k + (int)prmt(src, 0, 0x4499);
		return a + (t >> 16);
	}
	return m * k + (int)prmt(src, (uint32_t)a, 0x7699);
}

template <int C, bool SINGLE, bool SIGNED>
__device__ __forceinline__ void chain_tap(int (&a)[16], uint32_t frame, int k, int channels)
{
	const uint32_t ks = SIGNED ? (uint32_t)(k >> 31) : 0u;
	if (C == 2) {
		chain_word<SINGLE, SIGNED>(a[0], a[1], lds32(frame), k, ks);
	} else if (C == 4) {
		const uint2 v = lds64(frame);
		chain_word<SINGLE, SIGNED>(a[0], a[1], v.x, k, ks);
		chain_word<SINGLE, SIGNED>(a[2], a[3], v.y, k, ks);
	} else if (C == 8) {
		const uint4 v = lds128(frame);
		chain_word<SINGLE, SIGNED>(a[0], a[1], v.x, k, ks);
		chain_word<SINGLE, SIGNED>(a[2], a[3], v.y, k, ks);
		chain_word<SINGLE, SIGNED>(a[4], a[5], v.z, k, ks);
		chain_word<SINGLE, SIGNED>(a[6], a[7], v.w, k, ks);
	} else if (C == 6) {
		chain_word<SINGLE, SIGNED>(a[0], a[1], lds32(frame), k, ks);
		chain_word<SINGLE, SIGNED>(a[2], a[3], lds32(frame + 4), k, ks);
		chain_word<SINGLE, SIGNED>(a[4], a[5], lds32(frame + 8), k, ks);
	} else {
#pragma unroll
		for (int c = 0; c < 16; ++c)
			if (c < channels)
				a[c] = chain_scalar<SINGLE, SIGNED>(a[c], lds_s16(frame + 2 * c), k, ks);
	}
}

template <int C, bool SINGLE, bool SIGNED, bool CONSTOFF>
__device__ __forceinline__ void chain_pair(const crb_geometry &g, int (&a)[16], uint32_t w, uint32_t ci, uint32_t win, int channels)
{
	const uint2 kk = lds64(w);
	uint32_t o0, o1;
	if (CONSTOFF) { o0 = g.col_off16[ci]; o1 = g.col_off16[ci + 1]; }
	else { const uint2 oo = lds64(ci); o0 = oo.x; o1 = oo.y; }
	chain_tap<C, SINGLE, SIGNED>(a, win + o0, (int)kk.x, channels);
	chain_tap<C, SINGLE, SIGNED>(a, win + o1, (int)kk.y, channels);
}

/* One group of `count` columns (even): a chain accumulates in its own registers and is folded into `acc` at the end. */
template <int C, bool SINGLE, bool SIGNED, bool CONSTOFF>
__device__ __forceinline__ void chain_group(const crb_geometry &g, int (&acc)[16], uint32_t w, uint32_t ci, uint32_t win, uint32_t count, int channels)
{
	constexpr uint32_t CS = CONSTOFF ? 2u : 8u;
	const uint32_t pairs = count >> 1;
	if (SINGLE) {
		uint32_t i = 0;
		if (pairs & 1u) { chain_pair<C, true, SIGNED, CONSTOFF>(g, acc, w, ci, win, channels); w += 8; ci += CS; i = 1; }
#pragma unroll 1
		for (; i < pairs; i += 2, w += 16, ci += 2 * CS) {
			chain_pair<C, true, SIGNED, CONSTOFF>(g, acc, w, ci, win, channels);
			chain_pair<C, true, SIGNED, CONSTOFF>(g, acc, w + 8, ci + CS, win, channels);
		}
	} else {
		int ch[16];
#pragma unroll
		for (int c = 0; c < 16; ++c) ch[c] = 0;
		uint32_t i = 0;
		if (pairs & 1u) { chain_pair<C, false, SIGNED, CONSTOFF>(g, ch, w, ci, win, channels); w += 8; ci += CS; i = 1; }
#pragma unroll 1
		for (; i < pairs; i += 2, w += 16, ci += 2 * CS) {
			chain_pair<C, false, SIGNED, CONSTOFF>(g, ch, w, ci, win, channels);
			chain_pair<C, false, SIGNED, CONSTOFF>(g, ch, w + 8, ci + CS, win, channels);
		}
#pragma unroll
		for (int c = 0; c < 16; ++c) if (c < channels) acc[c] += ch[c] >> 16;
	}
}

template <int C, int FMT>
__device__ __forceinline__ void frame_chains(const crb_geometry &g, uint32_t t, uint32_t stage, uint32_t rows, unsigned char *outp, int channels, uint32_t lane_rot)
{
	const uint32_t fb = 2u * channels;
	const uint32_t e = ~t & 0xFFFFu;
	uint32_t r = (((e + g.delta) * g.step) >> 16) - g.ks0;
#pragma unroll
	for (uint32_t b = 0; b < CRB_MAX_BREAKS; ++b) r += (e >= g.breaks[b]);
	const uint32_t row = rows + r * g.row_words * 4;
	const uint32_t colinfo = rows + g.n_rows * g.row_words * 4;
	const uint32_t win = stage + (t >> 16) * fb;
	int accp[16], accn[16], outv[16];
#pragma unroll
	for (int c = 0; c < 16; ++c) accp[c] = accn[c] = 0;
	constexpr bool CAN_CONSTOFF = (C & 1) != 0;
	const bool constoff = CAN_CONSTOFF && g.const_offsets;
#pragma unroll 1
	for (uint32_t gi = 0; gi < g.n_groups; ++gi) {
		const uint32_t first = g.groups[gi][0], count = g.groups[gi][1];
		const uint32_t o = first * 4 + (lane_rot & g.group_rot[gi]);
#define CRB_CHAIN_GROUP(ACC, SINGLE, SIGNED) \
		if (CAN_CONSTOFF && constoff) chain_group<C, SINGLE, SIGNED, true>(g, ACC, row + o, first, win, count, channels); \
		else chain_group<C, SINGLE, SIGNED, false>(g, ACC, row + o, colinfo + o, win, count, channels);
		switch (g.group_kind[gi]) {
		case 0: CRB_CHAIN_GROUP(accp, false, false) break;
		case 1: CRB_CHAIN_GROUP(accp, true, false) break;
		case 2: CRB_CHAIN_GROUP(accn, false, false) break;
		case 3: CRB_CHAIN_GROUP(accn, true, false) break;
		case 4: CRB_CHAIN_GROUP(accp, false, true) break;
		default: CRB_CHAIN_GROUP(accp, true, true) break;
		}
#undef CRB_CHAIN_GROUP
	}
	const int recip_word = (int)lds32(row + g.n_cols * 4);
#define CRB_NORMALISE_ALL(MODE) \
	_Pragma("unroll") for (int c = 0; c < 16; ++c) if (c < channels) outv[c] = normalise(accp[c] - accn[c], recip_word, MODE);
	if (FMT == 2) {
#pragma unroll
		for (int c = 0; c < 16; ++c) if (c < channels) outv[c] = accp[c] - accn[c];
	} else if (g.norm_mode == 1) { CRB_NORMALISE_ALL(1)
	} else if (g.norm_mode == 3) { CRB_NORMALISE_ALL(3)
	} else if (g.norm_mode == 2) { CRB_NORMALISE_ALL(2)
	} else { CRB_NORMALISE_ALL(0) }
#undef CRB_NORMALISE_ALL
	store_frame<C, FMT>(outp, outv, channels, recip_of_row_word(recip_word, g.norm_mode));
}

/* One output frame of a slightly stretched kernel (TAPS = 6, 8, 10 or 12 taps, up to eight channels): the row holds the
   signed weights in tap order and the reciprocal word, fetched with 16-byte loads; the taps are unrolled with
   immediate frame offsets.  Every tap takes the signed big form (see tap_word): multiplicand sample << 16,
   bias sample ^ (k >> 31). */
template <int C, int FMT, int TAPS>
__device__ __forceinline__ void frame_sk(const crb_geometry &g, uint32_t t, uint32_t stage, uint32_t rows, unsigned char *outp, int channels)
{
	constexpr uint32_t RW = (TAPS + 1 + 3) & ~3u;
	constexpr int NC = C ? C : 8;        /* C == 0: channel count at run time (1..8), for the diagnostic format */
	const uint32_t fb = 2u * channels;
	const uint32_t e = ~t & 0xFFFFu;
	uint32_t r = (((e + g.delta) * g.step) >> 16) - g.ks0;
#pragma unroll
	for (uint32_t b = 0; b < CRB_MAX_BREAKS; ++b) r += (e >= g.breaks[b]);
	const uint32_t row = rows + r * 16;           /* planar table: plane q holds words 4q..4q+3 of every row (crb_dev_plan_upload) */
	const uint32_t plane = g.n_rows * 16;
	const uint32_t win = stage + (t >> 16) * fb;
	int w[RW];
#pragma unroll
	for (uint32_t q = 0; q < RW / 4; ++q) {
		const uint4 v = lds128(row + plane * q);
		w[4 * q] = (int)v.x; w[4 * q + 1] = (int)v.y; w[4 * q + 2] = (int)v.z; w[4 * q + 3] = (int)v.w;
	}
	int acc[16], outv[16];
#pragma unroll
	for (int c = 0; c < 16; ++c) acc[c] = 0;
#pragma unroll
	for (int j = 0; j < TAPS; ++j) {
		const int k = w[j];
		const uint32_t ks = (uint32_t)(k >> 31);
		if (C != 0 && C % 2 == 0 && TAPS <= 8) {
			/* even channel counts, few taps: one packed load and five ALU operations per channel pair; with more taps
			   the ALU pipe fills up first and two sign-extending loads with four operations win (measured on stereo:
			   6 taps 21 % faster packed, 10 and 12 taps 4-5 % faster split) */
#pragma unroll
			for (int c = 0; c < NC; c += 2) {
				const uint32_t wd = lds32(win + j * fb + 2 * c), wx = wd ^ ks;
				acc[c] = mac_trunc(acc[c], (int)prmt(wd, 0, 0x1044), k, prmt(wx, 0, 0x9910));
				acc[c + 1] = mac_trunc(acc[c + 1], (int)(wd & 0xFFFF0000u), k, (uint32_t)((int)wx >> 16));
			}
		} else {
#pragma unroll
			for (int c = 0; c < NC; ++c)
				if (c < channels) {
					const int m = lds_s16(win + j * fb + 2 * c);
					acc[c] = mac_trunc(acc[c], (int)prmt((uint32_t)m, 0, 0x1044), k, (uint32_t)m ^ ks);
				}
		}
	}
	const int recip_word = w[TAPS];
#define CRB_NORMALISE_ALL(MODE) \
	_Pragma("unroll") for (int c = 0; c < NC; ++c) outv[c] = normalise(acc[c], recip_word, MODE);
	if (FMT == 2) {
#pragma unroll
		for (int c = 0; c < NC; ++c) outv[c] = acc[c];
	} else if (g.norm_mode == 3) { CRB_NORMALISE_ALL(3)
	} else if (g.norm_mode == 2) { CRB_NORMALISE_ALL(2)
	} else if (g.norm_mode == 1) { CRB_NORMALISE_ALL(1)
	} else { CRB_NORMALISE_ALL(0) }
#undef CRB_NORMALISE_ALL
	store_frame<C, FMT>(outp, outv, channels, recip_of_row_word(recip_word, g.norm_mode));
}

/* K: 0 = general kernel, 1 = unstretched 5-column kernel, 6 / 8 / 10 / 12 = slightly stretched kernel with that many taps */
template <int C, int FMT, int K>
__global__ void __launch_bounds__(CRB_NT_K(C, K == 1) + 32, CRB_CTAS_K(C, K)) crb_tiled_kernel(const __grid_constant__ crb_kparams p)
{
	constexpr bool U5 = K == 1;
	constexpr int SK = K > 1 ? K : 0;
	constexpr uint32_t NT = CRB_NT_K(C, K == 1);     /* consumer threads */
	constexpr uint32_t FULL_TILE = CRB_FULL_TILE_K(C, K == 1);  /* tiles of exactly this many frames take the fully unrolled path */
	extern __shared__ __align__(128) unsigned char smem[];
	const crb_geometry &g = p.geo;
	uint64_t *full = (uint64_t *)smem;                                  /* [CRB_STAGES] */
	uint64_t *empty = full + CRB_STAGES;                                /* [CRB_STAGES] */
	crb_tile_info *infos = (crb_tile_info *)(smem + 128);               /* [2 * CRB_STAGES] */
	static_assert(128 + 2 * CRB_STAGES * sizeof(crb_tile_info) <= CRB_CTRL_BYTES && 16 * CRB_STAGES <= 128, "control block too small for the ring");
	unsigned char *rows_ptr = smem + CRB_CTRL_BYTES;
	const uint32_t rows_bytes = ((g.n_rows * g.row_words + g.colinfo_words) * 4 + 15u) & ~15u;
	unsigned char *stage0_ptr = rows_ptr + rows_bytes;
	const int channels = C ? C : (int)g.channels;
	const uint32_t tid = threadIdx.x;
	/* warp-uniform role index (broadcast so that the compiler may keep per-warp values in uniform registers) */
	const uint32_t warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);

	if (tid == 0) {
#pragma unroll
		for (int s = 0; s < CRB_STAGES; ++s) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], NT / 32);
		}
		mbar_fence_init();
	}
	/* the per-phase table stays resident for the life of the CTA */
	{
		const int4 *src = (const int4 *)p.rows;     /* the device copy is padded to a multiple of 16 bytes */
		int4 *dst = (int4 *)rows_ptr;
		for (uint32_t i = tid; i < rows_bytes / 16; i += NT + 32) dst[i] = src[i];
	}
	__syncthreads();

	if (warp == NT / 32) {
		/* ---- producer warp: feeds the ring, one lane per lockstep stream ---- */
		if (blockIdx.x < p.total_tiles) {
			const crb_device_job *jobs = job_table(p);
			uint32_t ji = find_job(p, blockIdx.x);
			crb_device_job job = jobs[ji];
			uint64_t next_base = ji + 1 < p.n_jobs ? jobs[ji + 1].tile_base : ~0ull;
			uint32_t it = 0;
#ifdef CRB_DEBUG_TIMING
			unsigned long long dbg_acc[3] = { 0, 0, 0 };
#endif
			for (uint64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
				const uint32_t s = it % CRB_STAGES;
				while (tile >= next_base) {    /* tiles are visited in increasing order: walk forward */
					++ji;
					job = jobs[ji];
					next_base = ji + 1 < p.n_jobs ? jobs[ji + 1].tile_base : ~0ull;
				}
				produce_tile(p, job, tile, stage0_ptr + s * g.stage_bytes, &infos[it % (2 * CRB_STAGES)], &full[s], it >= CRB_STAGES, &empty[s], ((it / CRB_STAGES) - 1) & 1, tid & 31u
#ifdef CRB_DEBUG_TIMING
					, dbg_acc
#endif
					);
			}
#ifdef CRB_DEBUG_TIMING
			if ((tid & 31u) == 0) { atomicAdd(&p.dbg[2], dbg_acc[0]); atomicAdd(&p.dbg[3], dbg_acc[1]); atomicAdd(&p.dbg[5], dbg_acc[2]); }
#endif
		}
		return;
	}

	/* ---- consumer warps: thread `tid` takes frames tid, tid + NT, ... of every tile ---- */
	const uint32_t fb_out = FMT == 1 ? channels * 2u : (channels + (FMT == 2)) * 4u;
	const uint32_t rows = smem_u32(rows_ptr);
	const uint32_t lane_rot = (U5 || SK) ? 0u : (((g.rot * (tid & 31u)) >> g.rot_shift) & g.rot_mask) * 8u;
	uint32_t it = 0;
#ifdef CRB_DEBUG_TIMING
	unsigned long long dbg_wait = 0, dbg_tiles = 0, dbg_work = 0;
#endif
	for (uint64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
		const uint32_t s = it % CRB_STAGES;
#ifdef CRB_DEBUG_TIMING
		const long long dbg_c0 = clock64();
#endif
		mbar_wait(&full[s], (it / CRB_STAGES) & 1);
#ifdef CRB_DEBUG_TIMING
		const long long dbg_c1 = clock64();
#endif

		const crb_tile_info &info = infos[it % (2 * CRB_STAGES)];
		const uint32_t increment = info.increment, n_frames = info.n_frames;
		const uint32_t t_step = NT * increment;

		/* (mono, s16: the full-tile path stores frame pairs as one 32-bit word and wants them 4-byte aligned) */
		bool fast = U5 && n_frames * info.n_streams == FULL_TILE;
		if (U5 && C == 1 && FMT == 1) {
#pragma unroll
			for (uint32_t q = 0; q < CRB_MAX_LOCKSTEP; ++q)
				if (q < info.n_streams && ((uintptr_t)info.out[q] & 3u)) fast = false;
		}
		if (fast) {
			/* full tile of the unstretched kernel (one, two or four lockstep streams) */
			if (info.n_streams == 1) u5_full_tile<C, FMT, 1, NT, FULL_TILE>(info, tid, rows, fb_out, channels);
			else if (info.n_streams == 2) u5_full_tile<C, FMT, 2, NT, FULL_TILE>(info, tid, rows, fb_out, channels);
			else u5_full_tile<C, FMT, 4, NT, FULL_TILE>(info, tid, rows, fb_out, channels);
		} else if (U5) {
			uint32_t tt = info.t0 + tid * increment;
			for (uint32_t j = tid; j < n_frames; j += NT, tt += t_step) {
				const uint4 r = u5_row(tt, rows);
				for (uint32_t q = 0; q < info.n_streams; ++q)
					frame_u5_row<C, FMT>(r, info.win[q] + (tt >> 16) * 2u * channels, info.out[q] + (size_t)j * fb_out, channels);
			}
		} else {
			/* thread tid takes frame (tid * lane_stride) mod NT of every NT-frame block (lane_stride is odd, so
			   this is a permutation): the plan picks the stride that spreads one load's lanes over the banks */
			const uint32_t f0 = SK ? tid : ((tid * g.lane_stride) & (NT - 1));
			const uint32_t stage = info.win[0];
			uint32_t tt = info.t0 + f0 * increment;
			unsigned char *o = info.out[0] + (size_t)f0 * fb_out;
			uint32_t j = f0;
			if (SK) {
				/* four frames per thread at a time: independent accumulator chains to overlap */
				for (; j + 3 * NT < n_frames; j += 4 * NT, tt += 4 * t_step, o += (size_t)4 * NT * fb_out) {
#pragma unroll
					for (uint32_t u = 0; u < 4; ++u)
						frame_sk<C, FMT, (SK ? SK : 6)>(g, tt + u * t_step, stage, rows, o + (size_t)u * NT * fb_out, channels);
				}
			}
			for (; j < n_frames; j += NT, tt += t_step, o += (size_t)NT * fb_out) {
				if (SK) frame_sk<C, FMT, (SK ? SK : 6)>(g, tt, stage, rows, o, channels);
				else if (g.chain_mode) frame_chains<C, FMT>(g, tt, stage, rows, o, channels, lane_rot);
				else frame_runs<C, FMT>(g, tt, stage, rows, o, channels, lane_rot);
			}
		}
		/* this warp is done with stage s */
		__syncwarp();
#ifdef CRB_DEBUG_TIMING
		const long long dbg_c2 = clock64();
		dbg_wait += (unsigned long long)(dbg_c1 - dbg_c0); dbg_tiles += 1ull; dbg_work += (unsigned long long)(dbg_c2 - dbg_c1);
#endif
		if ((tid & 31) == 0)
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
	}
#ifdef CRB_DEBUG_TIMING
	if ((tid & 31) == 0) {
		atomicAdd(&p.dbg[0], dbg_wait); atomicAdd(&p.dbg[1], dbg_tiles); atomicAdd(&p.dbg[4], dbg_work);
		/* buckets by (warp index in the CTA mod 4, hardware warp slot mod 4 = scheduler): wait, work, tiles */
		uint32_t hw; asm("mov.u32 %0, %%warpid;" : "=r"(hw));
		const uint32_t bucket = (warp & 3u) * 4u + (hw & 3u);
		atomicAdd(&p.dbg[8 + bucket], dbg_wait); atomicAdd(&p.dbg[24 + bucket], dbg_work); atomicAdd(&p.dbg[40 + bucket], dbg_tiles);
	}
#endif
}

typedef void (*crb_kernel_fn)(const crb_kparams);

/* crb_inst.cu: the instantiation of kernel kind K (0 general, 1 unstretched, 6 / 8 / 10 / 12 slightly stretched) for
   `channels` (0 = count at run time, 9..16) and output format `fmt`, or NULL; `*block` = threads per CTA it was compiled for.
   One function per (kind, channel range) translation unit. */
#define CRB_DECLARE_PICK(K, PART) extern "C" crb_kernel_fn crb_pick_k##K##_p##PART(unsigned channels, int fmt, unsigned *block);
CRB_DECLARE_PICK(0, 0) CRB_DECLARE_PICK(0, 1) CRB_DECLARE_PICK(1, 0) CRB_DECLARE_PICK(1, 1)
CRB_DECLARE_PICK(0, 2) CRB_DECLARE_PICK(0, 3) CRB_DECLARE_PICK(1, 2) CRB_DECLARE_PICK(1, 3)
CRB_DECLARE_PICK(6, 0) CRB_DECLARE_PICK(6, 1) CRB_DECLARE_PICK(8, 0) CRB_DECLARE_PICK(8, 1)
CRB_DECLARE_PICK(10, 0) CRB_DECLARE_PICK(10, 1) CRB_DECLARE_PICK(12, 0) CRB_DECLARE_PICK(12, 1)

#endif /* CRB_KERNELS_CUH */
