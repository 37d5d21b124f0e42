/*
 * crb_device.cu -- sm_100a kernels and CUDA-runtime glue of libclownresampler_b200.so.
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
#include <cuda_runtime.h>

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "crb_internal.h"

#define CRB_INLINE_JOBS 8
#define CRB_CTRL_BYTES 256

struct crb_kparams {
	crb_geometry geo;
	const int32_t *rows;
	const int32_t *table;
	const crb_device_job *jobs;
	uint32_t n_jobs;
	uint32_t out_format;   /* 0 s32, 1 s16 clamped, 2 s32 raw accumulators + reciprocal */
	uint64_t total_tiles;
	crb_device_job inline_jobs[CRB_INLINE_JOBS];
};

struct crb_tile_info {
	uint32_t t0;            /* (q - (ws0 - 1) * 65536) + 65535 for the tile's first frame */
	uint32_t n_frames;
	uint32_t lead_samples;  /* sample index inside the stage of input frame ws0 */
	uint32_t pad;
	unsigned char *out;     /* where the tile's first output frame goes */
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
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"CRB_WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra CRB_DONE_%=;\n"
		"bra CRB_WAIT_%=;\n"
		"CRB_DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* ------------------------------------------------------------------------------------------
 * the exact multiply-accumulate: acc + trunc_toward_zero(s * k / 65536), k >= 0
 *   S    = s << 16
 *   bias = 0 when s >= 0, 0xFFFFFFFF when s < 0
 * compiles to one IMAD.HI Rd, S, k, (acc:bias)
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int mac_trunc(int acc, int S, int k, uint32_t bias)
{
	const long long addend = (long long)(((unsigned long long)(uint32_t)acc << 32) | bias);
	return (int)(((long long)S * (long long)k + addend) >> 32);
}

/* PTX prmt in default mode: selector nibble bit 3 replicates the sign bit of the selected byte
   (the CUDA intrinsic __byte_perm masks that bit off, so it has to be inline PTX). */
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}

/* the two samples of a packed 32-bit word (lo = even channel, hi = odd channel):
   S = sample << 16, M = 0xFFFFFFFF for a negative sample else 0 */
__device__ __forceinline__ void unpack2(uint32_t w, int &S_lo, uint32_t &M_lo, int &S_hi, uint32_t &M_hi)
{
	S_lo = (int)prmt(w, 0, 0x1044);     /* bytes (0, 0, w.b0, w.b1) == w << 16, kept off the multiplier pipe */
	S_hi = (int)(w & 0xFFFF0000u);
	M_lo = prmt(w, 0, 0x9999);          /* byte 1 sign-replicated into all four bytes */
	M_hi = prmt(w, 0, 0xBBBB);          /* byte 3 sign-replicated */
}

/* one tap: all channels of input frame `frame` (a pointer to its first sample in shared memory) */
template <int C>
__device__ __forceinline__ void tap(int (&acc)[16], const int16_t *frame, int k, int channels)
{
	if (C == 1) {
		const uint32_t w = *(const uint16_t *)frame;
		acc[0] = mac_trunc(acc[0], (int)prmt(w, 0, 0x1044), k, prmt(w, 0, 0x9999));
	} else if (C == 2) {
		int s0, s1; uint32_t m0, m1;
		unpack2(*(const uint32_t *)frame, s0, m0, s1, m1);
		acc[0] = mac_trunc(acc[0], s0, k, m0);
		acc[1] = mac_trunc(acc[1], s1, k, m1);
	} else if (C == 4) {
		const uint2 v = *(const uint2 *)frame;
		int s0, s1; uint32_t m0, m1;
		unpack2(v.x, s0, m0, s1, m1);
		acc[0] = mac_trunc(acc[0], s0, k, m0); acc[1] = mac_trunc(acc[1], s1, k, m1);
		unpack2(v.y, s0, m0, s1, m1);
		acc[2] = mac_trunc(acc[2], s0, k, m0); acc[3] = mac_trunc(acc[3], s1, k, m1);
	} else if (C == 8) {
		const uint4 v = *(const uint4 *)frame;
		int s0, s1; uint32_t m0, m1;
		unpack2(v.x, s0, m0, s1, m1);
		acc[0] = mac_trunc(acc[0], s0, k, m0); acc[1] = mac_trunc(acc[1], s1, k, m1);
		unpack2(v.y, s0, m0, s1, m1);
		acc[2] = mac_trunc(acc[2], s0, k, m0); acc[3] = mac_trunc(acc[3], s1, k, m1);
		unpack2(v.z, s0, m0, s1, m1);
		acc[4] = mac_trunc(acc[4], s0, k, m0); acc[5] = mac_trunc(acc[5], s1, k, m1);
		unpack2(v.w, s0, m0, s1, m1);
		acc[6] = mac_trunc(acc[6], s0, k, m0); acc[7] = mac_trunc(acc[7], s1, k, m1);
	} else {
		/* any channel count 1..16: scalar 16-bit loads */
#pragma unroll
		for (int c = 0; c < 16; ++c)
			if (c < channels) {
				const int s = frame[c];
				acc[c] = mac_trunc(acc[c], s << 16, k, (uint32_t)(s >> 31));
			}
	}
}

/* out = trunc(acc * recip / 32768), H:1033.  Rows hold recip << 15 when every reciprocal of the
   plan is below 65536 (geo.recip_shift == 15; crb_plan.c proves |acc| < 2^29): then
   (acc << 2) * (recip << 15) = acc * recip * 2^17 and the same high-word trick as mac_trunc
   truncates toward zero in one IMAD.HI.  Otherwise the plain 64-bit form is used. */
__device__ __forceinline__ int normalise(int acc, int recip_row, uint32_t recip_shift)
{
	if (recip_shift == 15)
		return mac_trunc(0, acc << 2, recip_row, (uint32_t)(acc >> 31));
	const long long q = (long long)acc * (long long)recip_row;
	return (int)((q + ((q >> 63) & 32767)) >> 15);
}

__device__ __forceinline__ int clamp_s16(int v)
{
	return max(-0x7FFF, min(0x7FFF, v));   /* examples/low-level.c:74-77 */
}

template <int C, int FMT>
__device__ __forceinline__ void store_frame(unsigned char *out, uint32_t frame_index, const int (&v)[16], int channels, int recip)
{
	if (FMT == 2) {
		/* diagnostic format: un-normalised accumulators followed by the phase reciprocal */
		int *o = (int *)out + (size_t)frame_index * (channels + 1);
#pragma unroll
		for (int c = 0; c < 16; ++c) if (c < channels) o[c] = v[c];
		o[channels] = recip;
	} else if (FMT == 0) {
		int *o = (int *)out + (size_t)frame_index * channels;
		if (C == 2) { *(int2 *)o = make_int2(v[0], v[1]); }
		else if (C == 4) { *(int4 *)o = make_int4(v[0], v[1], v[2], v[3]); }
		else if (C == 8) { ((int4 *)o)[0] = make_int4(v[0], v[1], v[2], v[3]); ((int4 *)o)[1] = make_int4(v[4], v[5], v[6], v[7]); }
		else {
#pragma unroll
			for (int c = 0; c < 16; ++c) if (c < channels) o[c] = v[c];
		}
	} else {
		int16_t *o = (int16_t *)out + (size_t)frame_index * channels;
		if (C == 2) {
			*(uint32_t *)o = __byte_perm((uint32_t)clamp_s16(v[0]), (uint32_t)clamp_s16(v[1]), 0x5410);
		} else if (C == 4) {
			*(uint2 *)o = make_uint2(__byte_perm((uint32_t)clamp_s16(v[0]), (uint32_t)clamp_s16(v[1]), 0x5410),
			                         __byte_perm((uint32_t)clamp_s16(v[2]), (uint32_t)clamp_s16(v[3]), 0x5410));
		} else if (C == 8) {
			*(uint4 *)o = make_uint4(__byte_perm((uint32_t)clamp_s16(v[0]), (uint32_t)clamp_s16(v[1]), 0x5410),
			                         __byte_perm((uint32_t)clamp_s16(v[2]), (uint32_t)clamp_s16(v[3]), 0x5410),
			                         __byte_perm((uint32_t)clamp_s16(v[4]), (uint32_t)clamp_s16(v[5]), 0x5410),
			                         __byte_perm((uint32_t)clamp_s16(v[6]), (uint32_t)clamp_s16(v[7]), 0x5410));
		} else {
#pragma unroll
			for (int c = 0; c < 16; ++c) if (c < channels) o[c] = (int16_t)clamp_s16(v[c]);
		}
	}
}

__device__ __forceinline__ const crb_device_job *find_job(const crb_kparams &p, uint64_t tile)
{
	const crb_device_job *jobs = p.jobs ? p.jobs : p.inline_jobs;
	uint32_t lo = 0, hi = p.n_jobs;   /* last job with tile_base <= tile */
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (jobs[mid].tile_base <= tile) lo = mid; else hi = mid;
	}
	return jobs + lo;
}

/* Thread 0: describe tile `tile`, and start the bulk copy of its input window into `stage`. */
__device__ __forceinline__ void produce_tile(const crb_kparams &p, uint64_t tile, unsigned char *stage, crb_tile_info *info, uint64_t *bar)
{
	const crb_geometry &g = p.geo;
	const crb_device_job *job = find_job(p, tile);
	const uint64_t first = (tile - job->tile_base) * g.tile_out;
	const uint64_t left = job->n_out - first;
	const uint32_t n = left < g.tile_out ? (uint32_t)left : g.tile_out;
	const uint64_t q = job->q0 + (job->first_out + first) * (uint64_t)g.increment;
	const uint64_t ws0 = (q + 65535) >> 16;
	const uint64_t ws_last = (q + (uint64_t)(n - 1) * g.increment + 65535) >> 16;
	uint64_t end_frame = ws_last + g.taps_max;
	if (end_frame > job->in_frames) end_frame = job->in_frames;   /* columns past the buffer end are zero-weight */
	const uint32_t frame_bytes = 2 * g.channels;
	const uintptr_t a_first = (uintptr_t)job->in + ws0 * frame_bytes;
	const uintptr_t a_end = (uintptr_t)job->in + end_frame * frame_bytes;
	const uintptr_t a0 = a_first & ~(uintptr_t)15;
	uintptr_t a1 = (a_end + 15) & ~(uintptr_t)15;
	if (a1 <= a0) a1 = a0 + 16;
	const uint32_t bytes = (uint32_t)(a1 - a0);

	info->t0 = (uint32_t)(q - ((ws0 - 1) << 16)) + 65535u;
	info->n_frames = n;
	info->lead_samples = (uint32_t)(a_first - a0) >> 1;
	info->out = (unsigned char *)job->out + first * (p.out_format == 1 ? g.channels * (size_t)2 : (g.channels + (p.out_format == 2)) * (size_t)4);
	mbar_arrive_expect_tx(bar, bytes);
	tma_bulk_g2s(stage, (const void *)a0, bytes, bar);
}

/* ------------------------------------------------------------------------------------------
 * the tiled kernel
 *   C    : channels handled with packed vector loads (1, 2, 4, 8) or 0 = any count, scalar loads
 *   FMT  : 0 = s32 unclamped, 1 = s16 clamped
 *   U5   : unstretched 5-column kernel with compile-time signs + - + + - (step 1024, delta 0)
 * ------------------------------------------------------------------------------------------ */
template <int C, int FMT, bool U5>
__global__ void __launch_bounds__(CRB_THREADS, 2) crb_tiled_kernel(const __grid_constant__ crb_kparams p)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const crb_geometry &g = p.geo;
	uint64_t *bars = (uint64_t *)smem;                                  /* [2] */
	crb_tile_info *infos = (crb_tile_info *)(smem + 64);                /* [2] */
	int32_t *rows = (int32_t *)(smem + CRB_CTRL_BYTES);
	const uint32_t rows_bytes = g.n_rows * g.row_words * 4;
	unsigned char *stage0 = smem + CRB_CTRL_BYTES + rows_bytes;
	const int channels = C ? C : (int)g.channels;
	const uint32_t tid = threadIdx.x;

	if (tid == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		mbar_fence_init();
	}
	/* the per-phase table stays resident for the life of the CTA */
	{
		const int4 *src = (const int4 *)p.rows;
		int4 *dst = (int4 *)rows;
		for (uint32_t i = tid; i < rows_bytes / 16; i += CRB_THREADS) dst[i] = src[i];
	}
	__syncthreads();

	uint64_t tile = blockIdx.x;
	if (tid == 0 && tile < p.total_tiles)
		produce_tile(p, tile, stage0, &infos[0], &bars[0]);

	for (uint32_t it = 0; tile < p.total_tiles; ++it, tile += gridDim.x) {
		const uint32_t s = it & 1;
		/* every thread has left the previous use of stage s^1 (barrier at the loop end) */
		if (tid == 0 && tile + gridDim.x < p.total_tiles)
			produce_tile(p, tile + gridDim.x, stage0 + (s ^ 1) * g.stage_bytes, &infos[s ^ 1], &bars[s ^ 1]);
		mbar_wait(&bars[s], (it >> 1) & 1);

		const crb_tile_info info = infos[s];
		const int16_t *samples = (const int16_t *)(stage0 + s * g.stage_bytes) + info.lead_samples - channels;

		for (uint32_t j = tid; j < info.n_frames; j += CRB_THREADS) {
			const uint32_t t = info.t0 + j * g.increment;
			const uint32_t w1 = t >> 16;                 /* window start, frames after (ws0 - 1) */
			const uint32_t e = ~t & 0xFFFFu;             /* phase */
			const int16_t *win = samples + w1 * channels;
			int accp[16], accn[16], outv[16], recip;
#pragma unroll
			for (int c = 0; c < 16; ++c) accp[c] = accn[c] = 0;

			const int32_t *row;
			if (U5) {
				row = rows + (e >> 6) * 8;
				const int4 k03 = *(const int4 *)row;
				const int2 k4r = *(const int2 *)(row + 4);
				tap<C>(accp, win, k03.x, channels);
				tap<C>(accn, win + channels, k03.y, channels);
				tap<C>(accp, win + 2 * channels, k03.z, channels);
				tap<C>(accp, win + 3 * channels, k03.w, channels);
				tap<C>(accn, win + 4 * channels, k4r.x, channels);
				recip = k4r.y;
			} else {
				uint32_t r = (((e + g.delta) * g.step) >> 16) - g.ks0;
				for (uint32_t b = 0; b < g.n_breaks; ++b) r += (e >= g.breaks[b]);
				row = rows + r * g.row_words;
				for (uint32_t q = 0; q < g.n_runs; ++q) {
					const crb_run run = g.runs[q];
					const int32_t *w = row + run.col;
					const int16_t *f = win + run.off * channels;
					if (run.negative) {
#pragma unroll 4
						for (int i = 0; i < run.len; ++i) tap<C>(accn, f + i * channels, w[i], channels);
					} else {
#pragma unroll 4
						for (int i = 0; i < run.len; ++i) tap<C>(accp, f + i * channels, w[i], channels);
					}
				}
				recip = row[g.n_cols];
			}
#pragma unroll
			for (int c = 0; c < 16; ++c)
				if (c < channels) outv[c] = FMT == 2 ? accp[c] - accn[c] : normalise(accp[c] - accn[c], recip, g.recip_shift);
			store_frame<C, FMT>(info.out, j, outv, channels, recip >> g.recip_shift);
		}
		__syncthreads();
	}
}

/* ------------------------------------------------------------------------------------------
 * the direct kernel: one thread per output frame straight from global memory, evaluating the
 * reference's formulas (H:993-1033) with 64-bit integers and the original strided table.
 * Used when a tile's input window cannot fit shared memory (extreme down-sampling ratios),
 * and by the tests as an independent device-side cross-check of the tiled kernel.
 * ------------------------------------------------------------------------------------------ */
template <int FMT>
__global__ void __launch_bounds__(CRB_THREADS) crb_direct_kernel(const __grid_constant__ crb_kparams p)
{
	const crb_geometry &g = p.geo;
	for (uint64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
		const crb_device_job *job = find_job(p, tile);
		const uint64_t n = (tile - job->tile_base) * g.tile_out + threadIdx.x;
		if (n >= job->n_out) continue;
		const uint64_t q = job->q0 + (job->first_out + n) * (uint64_t)g.increment;
		const uint64_t ws = (q + 65535) >> 16;
		const uint32_t e = (uint32_t)((ws << 16) - q);
		const uint32_t frac = (uint32_t)((q - g.delta) & 0xFFFF);
		const uint32_t min_rel = (frac + g.delta + 65535u) >> 16;
		const uint32_t max_rel = (uint32_t)(((uint64_t)frac + g.radius_fx) >> 16);
		const uint32_t ntaps = g.radius_int + max_rel - min_rel;
		uint32_t kidx = (uint32_t)(((uint64_t)g.step * (e + g.delta)) >> 16);
		long long acc[16], sum = 0;
#pragma unroll
		for (int c = 0; c < 16; ++c) acc[c] = 0;
		const int16_t *f = job->in + ws * g.channels;
		for (uint32_t i = 0; i < ntaps; ++i, kidx += g.step, f += g.channels) {
			const long long k = p.table[kidx];
			sum += k;
#pragma unroll
			for (int c = 0; c < 16; ++c)
				if (c < (int)g.channels) acc[c] += (long long)f[c] * k / 65536;
		}
		const long long recip = 0x80000000ll / sum;
		int outv[16];
#pragma unroll
		for (int c = 0; c < 16; ++c)
			if (c < (int)g.channels) outv[c] = FMT == 2 ? (int)acc[c] : (int)(acc[c] * recip / 32768);
		unsigned char *out = (unsigned char *)job->out;
		if (FMT == 2) {
			int *o = (int *)out + n * (g.channels + 1);
			for (uint32_t c = 0; c < g.channels; ++c) o[c] = outv[c];
			o[g.channels] = (int)recip;
		} else if (FMT == 0) {
			int *o = (int *)out + n * g.channels;
			for (uint32_t c = 0; c < g.channels; ++c) o[c] = outv[c];
		} else {
			int16_t *o = (int16_t *)out + n * g.channels;
			for (uint32_t c = 0; c < g.channels; ++c) o[c] = (int16_t)clamp_s16(outv[c]);
		}
	}
}

/* ------------------------------------------------------------------------------------------
 * synthetic input and checksums (same integer hash as oracle/cr_oracle.c: cro_noise_sample)
 * ------------------------------------------------------------------------------------------ */
__host__ __device__ __forceinline__ uint32_t crb_mix32(uint32_t x)
{
	x ^= x >> 16; x *= 0x7FEB352Du;
	x ^= x >> 15; x *= 0x846CA68Bu;
	x ^= x >> 16;
	return x;
}

__global__ void crb_noise_kernel(int16_t *dst, uint32_t seed, uint32_t stream_id, uint64_t first_frame, uint64_t n_frames, uint32_t channels)
{
	const uint64_t total = n_frames * channels;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t frame = first_frame + i / channels;
		const uint32_t c = (uint32_t)(i % channels);
		uint32_t h = seed ^ 0x9E3779B9u;
		h = crb_mix32(h + stream_id * 0x85EBCA6Bu);
		h = crb_mix32(h + c * 0xC2B2AE35u);
		h = crb_mix32(h + (uint32_t)frame);
		h = crb_mix32(h + (uint32_t)(frame >> 32) * 0x27D4EB2Fu);
		dst[i] = (int16_t)(h >> 16);
	}
}

template <typename T>
__global__ void crb_checksum_kernel(const T *src, uint64_t words, unsigned long long *result)
{
	unsigned long long local = 0;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t v = (uint32_t)(int32_t)src[i];
		const uint32_t a = crb_mix32(v + (uint32_t)i * 0x9E3779B9u);
		const uint32_t b = crb_mix32(a ^ (uint32_t)(i >> 32));
		local += ((unsigned long long)a << 32 | b) ^ i;
	}
	for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o);
	if ((threadIdx.x & 31) == 0) atomicAdd(result, local);
}

/* ------------------------------------------------------------------------------------------
 * host glue
 * ------------------------------------------------------------------------------------------ */
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { crb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return -2; } } while (0)

static int g_device = -1;
static int g_sm_count = 0;
static uint32_t g_smem_optin = 0;

extern "C" int crb_dev_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" int crb_dev_init(int device)
{
	int n = 0, major = 0, v = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		crb_set_error("no usable CUDA device (%s); libclownresampler_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		return -1;
	}
	if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
	if (device >= n) { crb_set_error("device %d requested but only %d present", device, n); return -1; }
	CUDA_TRY(cudaSetDevice(device));
	CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
	if (major != 10) { crb_set_error("device %d is compute capability %d.x; this library contains sm_100a code only", device, major); return -1; }
	CUDA_TRY(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, device));
	CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
	g_smem_optin = (uint32_t)v;
	g_device = device;
	return 0;
}

extern "C" int crb_dev_current(void) { return g_device; }
extern "C" uint32_t crb_dev_smem_optin(void) { return g_smem_optin; }
extern "C" int crb_dev_sm_count(void) { return g_sm_count; }

extern "C" void *crb_dev_alloc(size_t bytes)
{
	void *p = NULL;
	cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
	if (e != cudaSuccess) { crb_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return NULL; }
	return p;
}
extern "C" void crb_dev_free(void *p) { if (p) cudaFree(p); }
extern "C" void *crb_dev_pinned_alloc(size_t bytes)
{
	void *p = NULL;
	cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 16);
	if (e != cudaSuccess) { crb_set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return NULL; }
	return p;
}
extern "C" void crb_dev_pinned_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int crb_dev_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
	return 0;
}
extern "C" int crb_dev_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	return 0;
}
extern "C" int crb_dev_sync(void *stream)
{
	CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
	return 0;
}
extern "C" void *crb_dev_stream_create(void)
{
	cudaStream_t s = NULL;
	if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return NULL; }
	return (void *)s;
}
extern "C" void crb_dev_stream_destroy(void *stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }
extern "C" void *crb_dev_event_create(void)
{
	cudaEvent_t e = NULL;
	if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return NULL; }
	return (void *)e;
}
extern "C" void crb_dev_event_destroy(void *event) { if (event) cudaEventDestroy((cudaEvent_t)event); }
extern "C" int crb_dev_event_record(void *event, void *stream) { CUDA_TRY(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream)); return 0; }
extern "C" int crb_dev_event_sync(void *event) { CUDA_TRY(cudaEventSynchronize((cudaEvent_t)event)); return 0; }
extern "C" int crb_dev_stream_wait_event(void *stream, void *event) { CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0)); return 0; }

extern "C" int crb_dev_plan_upload(struct ClownResamplerB200_Plan *plan)
{
	const size_t rows_bytes = (size_t)plan->geo.n_rows * plan->geo.row_words * 4;
	plan->dev_rows = crb_dev_alloc(rows_bytes);
	plan->dev_table = crb_dev_alloc(CRB_TABLE_SIZE * 4);
	if (!plan->dev_rows || !plan->dev_table) return -5;
	CUDA_TRY(cudaMemcpy(plan->dev_rows, plan->host_rows, rows_bytes, cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMemcpy(plan->dev_table, plan->host_table, CRB_TABLE_SIZE * 4, cudaMemcpyHostToDevice));
	plan->device = g_device;
	return 0;
}

extern "C" void crb_dev_plan_release(struct ClownResamplerB200_Plan *plan)
{
	crb_dev_free(plan->dev_rows); plan->dev_rows = NULL;
	crb_dev_free(plan->dev_table); plan->dev_table = NULL;
}

typedef void (*crb_kernel_fn)(const crb_kparams);

template <int C, int FMT>
static crb_kernel_fn pick_u5(bool u5)
{
	return u5 ? (crb_kernel_fn)crb_tiled_kernel<C, FMT, true> : (crb_kernel_fn)crb_tiled_kernel<C, FMT, false>;
}
template <int FMT>
static crb_kernel_fn pick_channels(unsigned channels, bool u5)
{
	switch (channels) {
	case 1: return pick_u5<1, FMT>(u5);
	case 2: return pick_u5<2, FMT>(u5);
	case 4: return pick_u5<4, FMT>(u5);
	case 8: return pick_u5<8, FMT>(u5);
	default: return pick_u5<0, FMT>(u5);
	}
}

extern "C" int crb_dev_launch(struct ClownResamplerB200_Plan *plan, const crb_device_job *jobs, size_t n_jobs,
	uint64_t total_tiles, int out_format, void *stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	crb_kparams p;
	crb_device_job *dev_jobs = NULL;
	if (total_tiles == 0 || n_jobs == 0) return 0;
	if (plan->device != g_device) { crb_set_error("plan was created on device %d but device %d is current", plan->device, g_device); return -4; }
	memset(&p, 0, sizeof p);
	p.geo = plan->geo;
	p.rows = (const int32_t *)plan->dev_rows;
	p.table = (const int32_t *)plan->dev_table;
	p.n_jobs = (uint32_t)n_jobs;
	p.out_format = (uint32_t)out_format;
	p.total_tiles = total_tiles;
	if (n_jobs <= CRB_INLINE_JOBS) {
		memcpy(p.inline_jobs, jobs, n_jobs * sizeof *jobs);
	} else {
		CUDA_TRY(cudaMallocAsync((void **)&dev_jobs, n_jobs * sizeof *jobs, stream));
		CUDA_TRY(cudaMemcpyAsync(dev_jobs, jobs, n_jobs * sizeof *jobs, cudaMemcpyHostToDevice, stream));
		p.jobs = dev_jobs;
	}

	if (plan->kernel_kind == 0) {
		const bool u5 = plan->geo.unstretched5 != 0;
		crb_kernel_fn fn = out_format == 1 ? pick_channels<1>(plan->geo.channels, u5)
		                 : out_format == 2 ? pick_u5<0, 2>(u5)
		                                   : pick_channels<0>(plan->geo.channels, u5);
		int per_sm = 0;
		CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes));
		CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, CRB_THREADS, plan->smem_bytes));
		if (per_sm < 1) { crb_set_error("tiled kernel does not fit an SM (%u bytes of shared memory)", plan->smem_bytes); return -2; }
		uint64_t grid = (uint64_t)g_sm_count * per_sm;
		if (grid > total_tiles) grid = total_tiles;
		void *args[] = { &p };
		CUDA_TRY(cudaLaunchKernel((const void *)fn, dim3((unsigned)grid), dim3(CRB_THREADS), args, plan->smem_bytes, stream));
	} else {
		uint64_t grid = (uint64_t)g_sm_count * 8;
		if (grid > total_tiles) grid = total_tiles;
		if (out_format == 1) crb_direct_kernel<1><<<(unsigned)grid, CRB_THREADS, 0, stream>>>(p);
		else if (out_format == 2) crb_direct_kernel<2><<<(unsigned)grid, CRB_THREADS, 0, stream>>>(p);
		else crb_direct_kernel<0><<<(unsigned)grid, CRB_THREADS, 0, stream>>>(p);
		CUDA_TRY(cudaGetLastError());
	}
	if (dev_jobs) CUDA_TRY(cudaFreeAsync(dev_jobs, stream));
	return 0;
}

extern "C" int crb_dev_fill_noise(int16_t *dst, uint32_t seed, uint32_t stream_id, uint64_t first_frame,
	uint64_t n_frames, uint32_t channels, void *stream)
{
	if (n_frames == 0) return 0;
	uint64_t blocks = (n_frames * channels + 255) / 256;
	if (blocks > (uint64_t)g_sm_count * 16) blocks = (uint64_t)g_sm_count * 16;
	crb_noise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, seed, stream_id, first_frame, n_frames, channels);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

extern "C" int crb_dev_checksum(const void *src, uint64_t words, int word_bytes, unsigned long long *result, void *stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	unsigned long long *d = NULL;
	CUDA_TRY(cudaMallocAsync((void **)&d, sizeof *d, stream));
	CUDA_TRY(cudaMemsetAsync(d, 0, sizeof *d, stream));
	uint64_t blocks = (words + 255) / 256;
	if (blocks > (uint64_t)g_sm_count * 16) blocks = (uint64_t)g_sm_count * 16;
	if (blocks == 0) blocks = 1;
	if (word_bytes == 2) crb_checksum_kernel<int16_t><<<(unsigned)blocks, 256, 0, stream>>>((const int16_t *)src, words, d);
	else crb_checksum_kernel<int32_t><<<(unsigned)blocks, 256, 0, stream>>>((const int32_t *)src, words, d);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(result, d, sizeof *d, cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	CUDA_TRY(cudaFreeAsync(d, stream));
	return 0;
}
