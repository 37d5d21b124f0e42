/*
 * clownresampler_b200.h -- additive C-ABI extensions of libclownresampler_b200.so.
 *
 * The drop-in API (clownresampler.h) keeps the reference's per-frame callback contract, which
 * is host-bound by construction.  These entry points expose the same hot path
 * (H:986-1035 + H:1058-1092, H = /root/reference/clownresampler.h) without callbacks:
 * plain pointers and sizes, no CUDA or torch types in any signature (streams travel as void*).
 * A reference maintainer binds these from C, or via ctypes/cgo/JNI as shown in INTEGRATION.md.
 *
 * Every function returns 0 on success and a negative CRB200_E_* code on failure unless stated
 * otherwise; ClownResamplerB200_GetLastError() returns the message of the calling thread's last
 * failure.  Nothing here falls back to the CPU.
 *
 * Threads: like the reference, the library is re-entrant per state.  The drop-in calls (clownresampler.h) and
 * ClownResamplerB200_ResampleHost borrow one of eight staging lanes (a CUDA stream and pinned + device buffers) for
 * the duration of a call, so calls on DIFFERENT states from different threads run concurrently (a ninth waits for a
 * lane); no lock is held while a kernel runs, while the GPU is awaited or while a user callback runs, so an output
 * callback may itself drive another resampler.  ClownResamplerB200_ResampleDevice takes no lock (plans are immutable
 * once created) and may be called concurrently, each caller on its own CUDA stream.  A
 * ClownResamplerB200_VoiceBatch must be used by one thread at a time.
 *
 * Devices: every plan, staging lane and voice batch lives on one device and remembers it; the entry points make that
 * device current for the calling thread and restore the caller's device before they return.  Calls without a handle
 * (the reference's C89 API, the allocation helpers) use the default device: the one last passed to
 * ClownResamplerB200_Init, else the CUDA device current in the calling thread at its first call.  One process can
 * drive all GPUs of a box: ClownResamplerB200_PlanCreateOnDevice / _DeviceAllocOn per device, one thread or CUDA
 * stream per device, or ClownResamplerB200_ResampleHostMulti, which does exactly that.
 *
 * Failures: the reference's API has no error channel.  When a drop-in call cannot compute (no device, a configuration the
 * reference itself cannot run, a CUDA error) it prints the reason to stderr, records it for
 * ClownResamplerB200_GetLastError() and returns cc_false WITHOUT consuming the input that produced no output (the state
 * advances exactly over the frames already delivered, as if the callback had asked to stop there).  CRB200_TRACE=1 makes a VoiceBatch print its per-tick phase times when destroyed;
 * CRB200_FORCE_DIRECT=1 selects the direct global-memory kernel for new plans and CRB200_NO_SMALL=1 keeps slightly
 * stretched kernels on the general kernel (test hooks).
 */
#ifndef CLOWNRESAMPLER_B200_H
#define CLOWNRESAMPLER_B200_H

#include "clownresampler.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
	CRB200_OK = 0,
	CRB200_E_NO_DEVICE = -1,     /* no usable CUDA device / driver */
	CRB200_E_CUDA = -2,          /* a CUDA runtime call or kernel failed */
	CRB200_E_CONFIG = -3,        /* configuration the reference would reject, crash on or overflow with */
	CRB200_E_ARGUMENT = -4,      /* bad pointer / size / alignment */
	CRB200_E_MEMORY = -5
};

enum {
	CRB200_OUT_S32 = 0,          /* unclamped 32-bit frames, what tests/test-low-level.c:43-49 writes */
	CRB200_OUT_S16_CLAMPED = 1,  /* clamp to [-0x7FFF, 0x7FFF] and narrow, examples/low-level.c:74-77 */
	CRB200_OUT_S32_RAW = 2       /* diagnostic: channels un-normalised accumulators (state before H:1025) followed by
	                                the frame's 17.15 reciprocal, channels + 1 words per frame; feeds the legacy-normaliser
	                                known-answer test against the reference's tests/test3 */
};

/* ---- lifecycle / errors --------------------------------------------------------------- */
int ClownResamplerB200_Init(int device);            /* optional: selects the default device (see "Devices" above) */
void ClownResamplerB200_Shutdown(void);             /* frees cached plans, staging buffers, streams */
const char *ClownResamplerB200_GetLastError(void);
int ClownResamplerB200_DeviceCount(void);

/* Diagnostics of the drop-in calls: kernel launches made by ClownResampler_LowLevel_Resample / HighLevel_* so far, and
   how many of their calls were served (at least partly) from frames a previous call had computed ahead of a callback
   that stopped it.  Such frames are reused only after the input they were computed from has been compared, byte for
   byte, with the input the new call presents. */
void ClownResamplerB200_GetCounters(unsigned long *dropin_kernel_launches, unsigned long *calls_served_from_kept_frames);

/* Plans built so far in this process (explicitly, or by the cache behind the calls without a plan handle).  A stream whose ratio is
   adjusted continuously must not build one per ratio: cached plans are keyed by kernel geometry, not by increment. */
unsigned long ClownResamplerB200_PlansBuilt(void);

/* ---- closed forms of the position generator (replaces the loop-carried H:1076-1078) ---- */
/* Frames H:1058-1092 would emit from this state over `total_input_frames` if never stopped. */
size_t ClownResamplerB200_CountOutputFrames(const ClownResampler_LowLevel_State *state, size_t total_input_frames);
/* Applies the end-of-call bookkeeping of H:1063-1067 (stopped == 0: input ran out after
   `frames_emitted` == Count frames) or H:1084-1088 (stopped != 0: the callback returned 0 on
   frame number `frames_emitted`). */
void ClownResamplerB200_AdvanceState(ClownResampler_LowLevel_State *state, size_t *total_input_frames,
	size_t frames_emitted, int stopped);

/* ---- plans: device-resident per-phase tap table for one (table, configuration, channels) ---- */
typedef struct ClownResamplerB200_Plan ClownResamplerB200_Plan;

typedef struct ClownResamplerB200_PlanInfo
{
	unsigned channels;
	unsigned long increment;
	unsigned phases;             /* rows of the per-phase table */
	unsigned taps_max;           /* widest tap window of any phase */
	unsigned columns;            /* multiply-accumulates per channel per frame the kernel issues */
	unsigned runs;               /* same-sign column runs */
	unsigned tile_output_frames; /* output frames per CTA tile */
	unsigned tile_input_frames;  /* input frames staged per tile (incl. halo) */
	unsigned smem_bytes;         /* dynamic shared memory per CTA */
	unsigned kernel_kind;        /* 0 = tiled shared-memory kernel, 1 = direct global-memory kernel */
	double mean_taps;            /* mean reference taps per frame over all 65536 fractions */
} ClownResamplerB200_PlanInfo;

/* `state` supplies lowest_level, channels and increment (as filled by ClownResampler_LowLevel_Init). */
ClownResamplerB200_Plan *ClownResamplerB200_PlanCreate(const ClownResampler_Precomputed *precomputed,
	const ClownResampler_LowLevel_State *state);
/* the same on a named device of this box (0 .. ClownResamplerB200_DeviceCount() - 1) instead of the default device */
ClownResamplerB200_Plan *ClownResamplerB200_PlanCreateOnDevice(const ClownResampler_Precomputed *precomputed,
	const ClownResampler_LowLevel_State *state, int device);
void ClownResamplerB200_PlanDestroy(ClownResamplerB200_Plan *plan);
int ClownResamplerB200_PlanGetInfo(const ClownResamplerB200_Plan *plan, ClownResamplerB200_PlanInfo *info);

/* ---- bulk resampling, no callbacks ------------------------------------------------------ */
/* One independent unit of work: a stream, or a contiguous output-time segment of one.
   `input` follows the H:725-733 padding contract (points at the leading R padding frames).
   The job emits output frames [first_output_frame, first_output_frame + output_frames) of the
   sequence H:1058-1092 would produce from (position_integer, position_fractional); frame
   first_output_frame is written at `output`.  Frames are `channels` interleaved samples. */
typedef struct ClownResamplerB200_Job
{
	const cc_s16l *input;
	void *output;
	size_t total_input_frames;       /* unpadded; used for bounds only */
	size_t position_integer;
	cc_u32f position_fractional;
	size_t first_output_frame;
	size_t output_frames;
} ClownResamplerB200_Job;

/* All pointers are DEVICE pointers; launches on `cuda_stream` (a cudaStream_t, NULL = default)
   and returns without synchronising.  `input` must be aligned to 4, 8, 4 and 16 bytes for 2, 4, 6 and 8 channels (the kernels of
   those counts use vector loads) and to 2 bytes for every other channel count. */
int ClownResamplerB200_ResampleDevice(ClownResamplerB200_Plan *plan, const ClownResamplerB200_Job *jobs,
	size_t job_count, int output_format, void *cuda_stream);

/* ---- planar device I/O (the format steps either side of the path) --------------------
   Decoders and mixers that keep one buffer per channel need not interleave first: a planar stream is `channels` mono streams
   that walk through the same positions, which the mono kernel runs as lockstep streams (one phase-row fetch per four channels).
   `plan` is a MONO plan of the stream's rates; every plane follows the H:725-733 padding contract on its own; output plane c
   receives the frames of channel c in `output_format` (s32, or s16 clamped to [-0x7FFF, 0x7FFF]). */
typedef struct ClownResamplerB200_PlanarJob
{
	const cc_s16l *const *input_planes;   /* channels device pointers, each at the start of its plane's leading padding */
	void *const *output_planes;           /* channels device pointers */
	size_t channels;
	size_t total_input_frames;
	size_t position_integer;
	cc_u32f position_fractional;
	size_t first_output_frame;
	size_t output_frames;
} ClownResamplerB200_PlanarJob;

int ClownResamplerB200_ResamplePlanarDevice(ClownResamplerB200_Plan *mono_plan, const ClownResamplerB200_PlanarJob *jobs,
	size_t job_count, int output_format, void *cuda_stream);

/* interleaved frames <-> planes on the device (16-bit or 32-bit words: word_bytes 2 or 4); all pointers are device pointers, the
   `planes` array itself is host memory.  Asynchronous on `cuda_stream`. */
int ClownResamplerB200_DeinterleaveDevice(const void *interleaved, void *const *planes, size_t frames, unsigned channels, int word_bytes, void *cuda_stream);
int ClownResamplerB200_InterleaveDevice(const void *const *planes, void *interleaved, size_t frames, unsigned channels, int word_bytes, void *cuda_stream);

/* All pointers are HOST pointers; stages through pinned memory with H2D, kernel and D2H
   overlapped in chunks, returns when `output` is complete. */
int ClownResamplerB200_ResampleHost(ClownResamplerB200_Plan *plan, const ClownResamplerB200_Job *jobs,
	size_t job_count, int output_format);

/* The same on several GPUs of this box from one call: `devices` lists `device_count` device indices.  `state` supplies the
   configuration, channel count and increment all jobs share (as for ClownResamplerB200_PlanCreate).  With at least as many
   jobs as devices the jobs are dealt to the devices in contiguous blocks; with fewer, every job is cut into device_count
   contiguous output-time segments (a segment needs only its slice of the input plus the kernel-radius halo,
   uploaded from the one buffer the segments share, so no data moves between GPUs).  One host thread per device; returns
   when every output is complete.  Formats: CRB200_OUT_S32, CRB200_OUT_S16_CLAMPED. */
int ClownResamplerB200_ResampleHostMulti(const ClownResampler_Precomputed *precomputed, const ClownResampler_LowLevel_State *state,
	const int *devices, size_t device_count, const ClownResamplerB200_Job *jobs, size_t job_count, int output_format);

/* Splits one stream into `segment_count` contiguous output ranges for multi-GPU / multi-job use
   (SURVEY.md 8e).  For segment `index` returns the output range, the slice of the padded input
   buffer it needs (in padded-buffer frames, halo included) and the start position RELATIVE to
   that slice, so that {input = padded + first*channels, position_*, first_output_frame = 0}
   is a valid job on a private copy of the slice. */
int ClownResamplerB200_SegmentStream(const ClownResampler_LowLevel_State *state, size_t total_input_frames,
	size_t segment_count, size_t index, size_t *first_output_frame, size_t *output_frames,
	size_t *first_padded_input_frame, size_t *padded_input_frames,
	size_t *position_integer, cc_u32f *position_fractional);

/* ---- batched streaming front end: many HighLevel-style voices per launch -------
   Every voice is a stream with the semantics of ClownResampler_HighLevel_Init / _Resample / _ResampleEnd
   (H:1101-1250): the stream is delayed by R frames, zero history before its start, R zero frames flushed at
   its end; the frames it emits are exactly the frames the reference's wrapper emits for the same input,
   however the input is chunked.  Instead of one GPU round trip per 4096-sample refill per voice, all voices
   advance together: Push() appends input (host memory, copied), Tick() produces up to `max_frames` frames for
   every voice with ONE input upload, ONE kernel launch per kernel geometry in use (normally one) and ONE output download.
   All voices of a batch share the channel count; their rates may differ (VoiceBatchAdjust). */
typedef struct ClownResamplerB200_VoiceBatch ClownResamplerB200_VoiceBatch;

ClownResamplerB200_VoiceBatch *ClownResamplerB200_VoiceBatchCreate(const ClownResampler_Precomputed *precomputed,
	size_t voices, cc_u8f channels, cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate);
void ClownResamplerB200_VoiceBatchDestroy(ClownResamplerB200_VoiceBatch *batch);
/* Appends `frames` interleaved input frames to voice `voice` (what its input callback would have delivered). */
int ClownResamplerB200_VoiceBatchPush(ClownResamplerB200_VoiceBatch *batch, size_t voice, const cc_s16l *input, size_t frames);
/* ClownResampler_HighLevel_Adjust (H:839) for one voice (pitch bend, or any other change of rates): the new rates apply from its
   next output frame on.  Voices are grouped by kernel geometry (H:632-638) tick by tick -- one launch per geometry in use, a plan
   per geometry built on first use -- so any rates are accepted that the reference's wrapper accepts: the new kernel radius must not
   exceed the one the batch was created with (H:1195) nor the wrapper's buffer (H:1202); else CRB200_E_CONFIG, voice unchanged. */
int ClownResamplerB200_VoiceBatchAdjust(ClownResamplerB200_VoiceBatch *batch, size_t voice,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate);
/* No more input for this voice: the remaining frames (incl. the R-frame flush of H:1216-1250) become available. */
int ClownResamplerB200_VoiceBatchEnd(ClownResamplerB200_VoiceBatch *batch, size_t voice);
/* One tick: voice v writes produced[v] <= max_frames frames at (char *)output + v * output_stride_bytes, in
   `output_format` (CRB200_OUT_S32 or CRB200_OUT_S16_CLAMPED).  A voice produces fewer than max_frames only when
   it has run out of pushed input (like H:1157 returning cc_true).  The rest of a voice's slot is unspecified after
   the call.  When `output` is pinned memory (ClownResamplerB200_PinnedAlloc) and the stride is a multiple of 16
   bytes, the download lands in it directly, without a staging copy. */
int ClownResamplerB200_VoiceBatchTick(ClownResamplerB200_VoiceBatch *batch, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced);
/* The same tick in two halves, for callers that have work of their own to overlap with the GPU (decoding the next tick's input,
   mixing the previous tick's output): TickBegin gathers the input, fills `produced` and queues upload, kernels and download on
   the batch's stream; TickEnd waits for them and completes `output`.  Between the two, `output` and `produced` belong to the
   library and the batch accepts Push and End only.  VoiceBatchTick is TickBegin followed by TickEnd. */
int ClownResamplerB200_VoiceBatchTickBegin(ClownResamplerB200_VoiceBatch *batch, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced);
int ClownResamplerB200_VoiceBatchTickEnd(ClownResamplerB200_VoiceBatch *batch);

/* ---- device helpers for C callers that do not link the CUDA runtime themselves ---------- */
void *ClownResamplerB200_DeviceAlloc(size_t bytes);                 /* on the default device */
void *ClownResamplerB200_DeviceAllocOn(int device, size_t bytes);
void ClownResamplerB200_DeviceFree(void *device_pointer);
void *ClownResamplerB200_PinnedAlloc(size_t bytes);
void ClownResamplerB200_PinnedFree(void *host_pointer);
int ClownResamplerB200_CopyToDevice(void *device_dst, const void *host_src, size_t bytes);
int ClownResamplerB200_CopyToHost(void *host_dst, const void *device_src, size_t bytes);
int ClownResamplerB200_Synchronize(void *cuda_stream);              /* NULL: the default stream of the default device */
int ClownResamplerB200_SynchronizeOn(int device, void *cuda_stream);

/* Deterministic synthetic s16 input generated on the device (counter-based integer hash of
   (seed, stream, channel, frame); the same generator exists on the host side of the tests so
   any window can be regenerated).  Writes n_frames * channels samples. */
int ClownResamplerB200_FillNoiseDevice(cc_s16l *device_dst, unsigned seed, unsigned stream,
	size_t first_frame, size_t n_frames, unsigned channels, void *cuda_stream);
/* Order-dependent 64-bit checksum of a device buffer of 16-bit or 32-bit words
   (sum over i of mix(word_i, i)); used for checksum-of-checksums parity at full size. */
int ClownResamplerB200_ChecksumDevice(const void *device_src, size_t words, int word_bytes,
	unsigned long *host_result, void *cuda_stream);

/* Test hook: builds the per-phase table on the host only (no device) and serialises it; see
   clownresampler_b200/csrc/crb_api.c for the word layout.  Returns geometry words or < 0. */
int ClownResamplerB200_DebugBuildPlanHost(const ClownResampler_Precomputed *precomputed,
	const ClownResampler_LowLevel_State *state, unsigned smem_budget_bytes,
	unsigned *geometry_words, size_t geometry_capacity, int *rows, size_t rows_capacity);

#ifdef __cplusplus
}
#endif

#endif /* CLOWNRESAMPLER_B200_H */
