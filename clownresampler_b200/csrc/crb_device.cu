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
#include "crb_kernels.cuh"

#include <mutex>

#ifdef CRB_DEBUG_TIMING
static unsigned long long *g_dbg;
extern "C" __attribute__((visibility("default"))) void ClownResamplerB200_DebugTiming(unsigned long long *out, int reset)
{
	cudaDeviceSynchronize();
	/* words 0..7: totals; 8..39: wait cycles per consumer warp index; 40..71: work cycles per consumer warp index */
	if (!g_dbg) { memset(out, 0, 576); return; }
	cudaMemcpy(out, g_dbg, 576, cudaMemcpyDeviceToHost);
	if (reset) cudaMemset(g_dbg, 0, 576);
}
#endif

/* ------------------------------------------------------------------------------------------
 * the direct kernel: one thread per output frame straight from global memory, evaluating the
 * reference's formulas (H:993-1033) with 64-bit integers and the original strided table.
 * Used when a tile's input window cannot fit shared memory (extreme down-sampling ratios),
 * and by the tests as an independent device-side cross-check of the tiled kernel.
 * ------------------------------------------------------------------------------------------ */
template <int FMT>
__global__ void __launch_bounds__(CRB_DIRECT_THREADS) crb_direct_kernel(const __grid_constant__ crb_kparams p)
{
	const crb_geometry &g = p.geo;
	for (uint64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
		const crb_device_job *job = job_table(p) + find_job(p, tile);
		const uint64_t n = (tile - job->tile_base) * g.tile_out + threadIdx.x;
		if (n >= job->n_out) continue;
		const uint64_t q = job->q0 + (job->first_out + n) * (job->increment ? job->increment : (uint64_t)g.increment);
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

/* ------------------------------------------------------------------------------------------
 * format steps either side of the path: planar <-> interleaved frames on the device.
 * HBM-bound copies: every thread moves 16 bytes of one plane (8 samples of 16 bits or 4 of 32) and scatters /
 * gathers them to `channels` interleaved frames; the plane side is fully coalesced, the interleaved side is
 * coalesced across the channel loop through L2 (the frames of one thread are `channels` words apart).
 * ------------------------------------------------------------------------------------------ */
struct crb_planes { void *plane[CRB_MAX_CHANNELS]; };

template <typename T>
__global__ void crb_deinterleave_kernel(const T *__restrict__ src, const __grid_constant__ crb_planes dst, uint64_t frames, uint32_t channels)
{
	/* one thread per frame: reads the frame's `channels` samples (consecutive), writes one sample to every plane (coalesced) */
	for (uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; f < frames; f += (uint64_t)gridDim.x * blockDim.x) {
		const T *in = src + f * channels;
		for (uint32_t c = 0; c < channels; ++c) ((T *)dst.plane[c])[f] = in[c];
	}
}

template <typename T>
__global__ void crb_interleave_kernel(const __grid_constant__ crb_planes src, T *__restrict__ dst, uint64_t frames, uint32_t channels)
{
	for (uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; f < frames; f += (uint64_t)gridDim.x * blockDim.x) {
		T *out = dst + f * channels;
		for (uint32_t c = 0; c < channels; ++c) out[c] = ((const T *)src.plane[c])[f];
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

/* Per-device context.  The library holds no "current device" of its own: every plan, staging lane and voice batch
   remembers the device it lives on, entry points make that device current for the calling thread (crb_dev_push) and
   put the caller's device back when they return (crb_dev_pop).  Calls without a handle (the reference's C89 API,
   allocation helpers) use the DEFAULT device: the one last passed to ClownResamplerB200_Init, else the CUDA device
   current in the calling thread at its first call. */
#define CRB_MAX_DEVICES 64
static struct crb_devctx { int ready; int sm_count; uint32_t smem_optin; } g_ctx[CRB_MAX_DEVICES];
static int g_default_device = -1;
static std::mutex g_ctx_lock;

extern "C" int crb_dev_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

/* makes `device` usable (device < 0: the default device) and returns its index, or a negative error */
extern "C" int crb_dev_init(int device, int make_default)
{
	int n = 0, major = 0, v = 0, sms = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		crb_set_error("no usable CUDA device (%s); libclownresampler_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		return -1;
	}
	std::lock_guard<std::mutex> guard(g_ctx_lock);
	if (device < 0) device = g_default_device;
	if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; } }
	if (device >= n || device >= CRB_MAX_DEVICES) { crb_set_error("device %d requested but only %d present", device, n); return -1; }
	if (!g_ctx[device].ready) {
		int prev = -1;
		cudaGetDevice(&prev);
		CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
		if (major != 10) { crb_set_error("device %d is compute capability %d.x; this library contains sm_100a code only", device, major); return -1; }
		CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
		CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
		CUDA_TRY(cudaSetDevice(device));
		{
			/* job tables of big batches come from the stream-ordered allocator: keep its pool instead of
			   returning memory to the driver at every synchronisation (that costs about a millisecond per launch) */
			cudaMemPool_t pool;
			if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
				unsigned long long keep = ~0ull;
				cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
			}
			cudaGetLastError();
		}
		if (prev >= 0 && prev != device) cudaSetDevice(prev);
		g_ctx[device].sm_count = sms;
		g_ctx[device].smem_optin = (uint32_t)v;
		g_ctx[device].ready = 1;
	}
	if (make_default || g_default_device < 0) g_default_device = device;
	return device;
}

extern "C" int crb_dev_default(void) { return g_default_device; }
extern "C" uint32_t crb_dev_smem_optin(int device) { return device >= 0 && device < CRB_MAX_DEVICES ? g_ctx[device].smem_optin : 0; }
extern "C" int crb_dev_sm_count(int device) { return device >= 0 && device < CRB_MAX_DEVICES ? g_ctx[device].sm_count : 0; }

/* makes `device` current for the calling thread; returns the device that was current (pass it to crb_dev_pop), or -1 */
extern "C" int crb_dev_push(int device)
{
	int prev = -1;
	if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
	if (prev != device && cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return -1; }
	return prev;
}
extern "C" void crb_dev_pop(int previous)
{
	int cur = -1;
	if (previous < 0) return;
	if (cudaGetDevice(&cur) == cudaSuccess && cur != previous) cudaSetDevice(previous);
}

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
/* the device a device pointer lives on, or -1 (not a device pointer) */
extern "C" int crb_dev_of_pointer(const void *p)
{
	cudaPointerAttributes a;
	if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
	return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}
/* 1 when [p, p + bytes) is page-locked host memory the copy engines can address directly */
extern "C" int crb_dev_is_pinned(const void *p, size_t bytes)
{
	cudaPointerAttributes a, b;
	if (!p || bytes == 0) return 0;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess || cudaPointerGetAttributes(&b, (const char *)p + bytes - 1) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return a.type == cudaMemoryTypeHost && b.type == cudaMemoryTypeHost;
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
	const size_t rows_bytes = ((size_t)plan->geo.n_rows * plan->geo.row_words + plan->geo.colinfo_words) * 4;
	plan->dev_rows = crb_dev_alloc(rows_bytes + 16);   /* the kernel copies it in whole 16-byte words */
	plan->dev_table = crb_dev_alloc(CRB_TABLE_SIZE * 4);
	if (!plan->dev_rows || !plan->dev_table) return -5;
	if (plan->geo.small_taps) {
		/* slightly stretched kernel: the device table is PLANAR -- plane q holds words 4q..4q+3 of every row, 16 bytes
		   per row -- so that the 16-byte loads of lanes on different rows spread over all eight bank groups
		   (row-major rows of 32 or 64 bytes would alias on two or four of them: measured 60 % conflict wavefronts) */
		const uint32_t n = plan->geo.n_rows, rw = plan->geo.row_words;
		int32_t *planar = (int32_t *)malloc(rows_bytes);
		if (!planar) return -5;
		for (uint32_t r = 0; r < n; ++r)
			for (uint32_t i = 0; i < rw; ++i)
				planar[(size_t)(i / 4) * n * 4 + (size_t)r * 4 + (i & 3u)] = plan->host_rows[(size_t)r * rw + i];
		cudaError_t e = cudaMemcpy(plan->dev_rows, planar, rows_bytes, cudaMemcpyHostToDevice);
		free(planar);
		CUDA_TRY(e);
	} else {
		CUDA_TRY(cudaMemcpy(plan->dev_rows, plan->host_rows, rows_bytes, cudaMemcpyHostToDevice));
	}
	CUDA_TRY(cudaMemcpy(plan->dev_table, plan->host_table, CRB_TABLE_SIZE * 4, cudaMemcpyHostToDevice));
	if (cudaGetDevice(&plan->device) != cudaSuccess) { cudaGetLastError(); plan->device = -1; }
	return 0;
}

extern "C" void crb_dev_plan_release(struct ClownResamplerB200_Plan *plan)
{
	crb_dev_free(plan->dev_rows); plan->dev_rows = NULL;
	crb_dev_free(plan->dev_table); plan->dev_table = NULL;
}

/* the instantiation for (channels, format, kernel kind) and the block size it was compiled for (crb_inst.cu);
   kind: 0 general, 1 unstretched, 6 / 8 / 10 / 12 slightly stretched (1..8 channels; the diagnostic format through C == 0) */
static crb_kernel_fn pick_kernel(unsigned channels, int fmt, unsigned kind, unsigned *block)
{
	/* the diagnostic format, and any count without an instantiation of its own, take the run-time channel count (C = 0) */
	const unsigned c = (fmt == 2 || channels > 16) ? 0 : channels;
	const int part = c >= 13 ? 3 : c >= 9 ? 2 : c >= 5 ? 1 : 0;
	switch (kind) {
	case 0: return part == 3 ? crb_pick_k0_p3(c, fmt, block) : part == 2 ? crb_pick_k0_p2(c, fmt, block) : part == 1 ? crb_pick_k0_p1(c, fmt, block) : crb_pick_k0_p0(c, fmt, block);
	case 1: return part == 3 ? crb_pick_k1_p3(c, fmt, block) : part == 2 ? crb_pick_k1_p2(c, fmt, block) : part == 1 ? crb_pick_k1_p1(c, fmt, block) : crb_pick_k1_p0(c, fmt, block);
	case 6: return part == 1 ? crb_pick_k6_p1(c, fmt, block) : crb_pick_k6_p0(c, fmt, block);
	case 8: return part == 1 ? crb_pick_k8_p1(c, fmt, block) : crb_pick_k8_p0(c, fmt, block);
	case 10: return part == 1 ? crb_pick_k10_p1(c, fmt, block) : crb_pick_k10_p0(c, fmt, block);
	case 12: return part == 1 ? crb_pick_k12_p1(c, fmt, block) : crb_pick_k12_p0(c, fmt, block);
	}
	return (crb_kernel_fn)NULL;
}

static int launch_jobs(struct ClownResamplerB200_Plan *plan, const crb_device_job *jobs, const crb_device_job *resident_jobs, size_t n_jobs,
	uint64_t total_tiles, int out_format, void *stream_);

extern "C" int crb_dev_launch(struct ClownResamplerB200_Plan *plan, const crb_device_job *jobs, size_t n_jobs,
	uint64_t total_tiles, int out_format, void *stream)
{
	return launch_jobs(plan, jobs, NULL, n_jobs, total_tiles, out_format, stream);
}

extern "C" int crb_dev_launch_resident(struct ClownResamplerB200_Plan *plan, const crb_device_job *device_jobs, size_t n_jobs,
	uint64_t total_tiles, int out_format, void *stream)
{
	return launch_jobs(plan, NULL, device_jobs, n_jobs, total_tiles, out_format, stream);
}

static int launch_jobs(struct ClownResamplerB200_Plan *plan, const crb_device_job *jobs, const crb_device_job *resident_jobs, size_t n_jobs,
	uint64_t total_tiles, int out_format, void *stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	crb_kparams p;
	crb_device_job *dev_jobs = NULL;
	if (total_tiles == 0 || n_jobs == 0) return 0;
	const int sm_count = crb_dev_sm_count(plan->device);
	{
		/* launches go to the CUDA device current in the calling thread; the public entry points have made it the plan's */
		int cur = -1;
		if (cudaGetDevice(&cur) != cudaSuccess || cur != plan->device) {
			cudaGetLastError();
			crb_set_error("plan lives on device %d but device %d is current in this thread", plan->device, cur);
			return -4;
		}
	}
	memset(&p, 0, sizeof p);
	p.geo = plan->geo;
	p.rows = (const int32_t *)plan->dev_rows;
	p.table = (const int32_t *)plan->dev_table;
	p.n_jobs = (uint32_t)n_jobs;
	p.out_format = (uint32_t)out_format;
	p.total_tiles = total_tiles;
#ifdef CRB_DEBUG_TIMING
	if (!g_dbg) { cudaMalloc((void **)&g_dbg, 576); cudaMemset(g_dbg, 0, 576); }
	p.dbg = g_dbg;
#endif
	if (resident_jobs) {
		p.jobs = resident_jobs;
	} else if (n_jobs <= CRB_INLINE_JOBS) {
		memcpy(p.inline_jobs, jobs, n_jobs * sizeof *jobs);
	} else {
		CUDA_TRY(cudaMallocAsync((void **)&dev_jobs, n_jobs * sizeof *jobs, stream));
		CUDA_TRY(cudaMemcpyAsync(dev_jobs, jobs, n_jobs * sizeof *jobs, cudaMemcpyHostToDevice, stream));
		p.jobs = dev_jobs;
	}

	if (plan->kernel_kind == 0) {
		const unsigned kind = plan->geo.unstretched5 ? 1u : plan->geo.small_taps;
		unsigned block = 0;
		crb_kernel_fn fn = pick_kernel(plan->geo.channels, out_format, kind, &block);
		if (!fn) { crb_set_error("no kernel instantiation for this plan (kind %u, %u channels, format %d)", kind, plan->geo.channels, out_format); return -2; }
		int per_sm;
		{
			/* First launch of this plan in this format: opt in to the shared memory and size the persistent grid.  The
			   opt-in is a property of the FUNCTION on a device, shared by every plan that uses it: only ever raise it.
			   The plan caches the result per output format; both tables are guarded by one mutex. */
			static struct { const void *fn; int device; uint32_t bytes; } optin[128];
			static int n_optin;
			static std::mutex optin_lock;
			std::lock_guard<std::mutex> guard(optin_lock);
			per_sm = plan->blocks_per_sm[out_format];
			if (plan->launch_fn[out_format] != (const void *)fn) {
				int i = 0;
				while (i < n_optin && !(optin[i].fn == (const void *)fn && optin[i].device == plan->device)) ++i;
				if (i == n_optin && n_optin < 128) { optin[n_optin].fn = (const void *)fn; optin[n_optin].device = plan->device; optin[n_optin].bytes = 0; ++n_optin; }
				if (i == 128 || optin[i].bytes < plan->smem_bytes) {
					CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes));
					if (i < 128) optin[i].bytes = plan->smem_bytes;
				}
				CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, (int)block, plan->smem_bytes));
				if (per_sm < 1) { crb_set_error("tiled kernel does not fit an SM (%u bytes of shared memory)", plan->smem_bytes); return -2; }
				plan->launch_fn[out_format] = (const void *)fn;
				plan->blocks_per_sm[out_format] = per_sm;
			}
		}
		uint64_t grid = (uint64_t)sm_count * per_sm;
		if (grid > total_tiles) grid = total_tiles;
		void *args[] = { &p };
		CUDA_TRY(cudaLaunchKernel((const void *)fn, dim3((unsigned)grid), dim3(block), args, plan->smem_bytes, stream));
	} else {
		uint64_t grid = (uint64_t)sm_count * 8;
		if (grid > total_tiles) grid = total_tiles;
		if (out_format == 1) crb_direct_kernel<1><<<(unsigned)grid, CRB_DIRECT_THREADS, 0, stream>>>(p);
		else if (out_format == 2) crb_direct_kernel<2><<<(unsigned)grid, CRB_DIRECT_THREADS, 0, stream>>>(p);
		else crb_direct_kernel<0><<<(unsigned)grid, CRB_DIRECT_THREADS, 0, stream>>>(p);
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
	if (blocks > (uint64_t)148 * 16) blocks = (uint64_t)148 * 16;
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
	if (blocks > (uint64_t)148 * 16) blocks = (uint64_t)148 * 16;
	if (blocks == 0) blocks = 1;
	if (word_bytes == 2) crb_checksum_kernel<int16_t><<<(unsigned)blocks, 256, 0, stream>>>((const int16_t *)src, words, d);
	else crb_checksum_kernel<int32_t><<<(unsigned)blocks, 256, 0, stream>>>((const int32_t *)src, words, d);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(result, d, sizeof *d, cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	CUDA_TRY(cudaFreeAsync(d, stream));
	return 0;
}

extern "C" int crb_dev_interleave(void *const *planes, void *interleaved, uint64_t frames, uint32_t channels, int word_bytes, int to_planes, void *stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	crb_planes p;
	if (frames == 0) return 0;
	memset(&p, 0, sizeof p);
	for (uint32_t c = 0; c < channels; ++c) p.plane[c] = planes[c];
	uint64_t blocks = (frames + 255) / 256;
	if (blocks > (uint64_t)148 * 32) blocks = (uint64_t)148 * 32;
	if (word_bytes == 2) {
		if (to_planes) crb_deinterleave_kernel<int16_t><<<(unsigned)blocks, 256, 0, stream>>>((const int16_t *)interleaved, p, frames, channels);
		else crb_interleave_kernel<int16_t><<<(unsigned)blocks, 256, 0, stream>>>(p, (int16_t *)interleaved, frames, channels);
	} else {
		if (to_planes) crb_deinterleave_kernel<int32_t><<<(unsigned)blocks, 256, 0, stream>>>((const int32_t *)interleaved, p, frames, channels);
		else crb_interleave_kernel<int32_t><<<(unsigned)blocks, 256, 0, stream>>>(p, (int32_t *)interleaved, frames, channels);
	}
	CUDA_TRY(cudaGetLastError());
	return 0;
}
