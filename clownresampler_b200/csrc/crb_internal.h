/*
 * crb_internal.h -- structures shared by the plain-C host layer (crb_api.c, crb_plan.c) and the
 * CUDA layer (crb_device.cu).  Not installed; the public surface is the two headers under include/.
 *
 * H = /root/reference/clownresampler.h.
 */
#ifndef CRB_INTERNAL_H
#define CRB_INTERNAL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRB_TABLE_SIZE 6144          /* H:629 */
#define CRB_FX_ONE 65536u            /* H:620 */
#define CRB_MAX_CHANNELS 16          /* H:458-460 */
/* Consumer threads per CTA of the tiled kernel (a power of two; one more warp produces), the CTAs per SM each
   instantiation is compiled and sized for, and the frames of a "full" tile (fully unrolled path: 16 per thread).
   Measured on B200: 1/2/4 channels run best as 2 CTAs x 512 threads with 8192-frame tiles (one table copy and one
   producer warp per 16 consumer warps); 8 channels need 71 registers: 3 CTAs x 256 threads; other counts 2 x 256. */
#define CRB_NT(channels) (((channels) == 1 || (channels) == 2 || (channels) == 4) ? 512 : 256)
#define CRB_CTAS(channels) (((channels) == 1 || (channels) == 2 || (channels) == 4) ? 2 : (channels) == 8 ? 3 : 2)
/* ... per kernel kind (0 general, 1 unstretched, 6..12 slightly stretched): the 8-channel general kernel takes two CTAs with 512-frame
   tiles rather than three with 256 (two frames per thread and tile instead of one; measured 7.21 -> 6.99 ms on config 3) */
#define CRB_CTAS_K(channels, kind) (((channels) == 8 && (kind) == 0) ? 2 : CRB_CTAS(channels))
/* The stereo unstretched kernel fits 40 registers with the two-instruction multiply-accumulate: 22 consumer warps per CTA instead
   of 16 (measured: 20 warps 1.5 % faster than 16, 22 another 1 %; 18 the same as 16, 21 slower -- an odd count loads the four
   schedulers unevenly --, 24 slower; the mono kernel with its frame pairs spills at 40 registers and is slower).  The thread
   count of an unstretched kernel need not be a power of two. */
#ifndef CRB_NT_STEREO_U5
#define CRB_NT_STEREO_U5 704
#endif
#define CRB_NT_K(channels, unstretched) (((unstretched) && (channels) == 2) ? CRB_NT_STEREO_U5 : CRB_NT(channels))
#ifndef CRB_FRAMES_PER_THREAD
#define CRB_FRAMES_PER_THREAD 16
#endif
#define CRB_FULL_TILE_K(channels, unstretched) (CRB_FRAMES_PER_THREAD * CRB_NT_K(channels, unstretched))
#define CRB_MAX_RUNS 24
#define CRB_MAX_BREAKS 4
#define CRB_CONST_COLS 384          /* columns whose frame offsets fit the kernel parameters */
#define CRB_GROUPS 12             /* column groups of the general kernel.  IMAD.HI form: six, (positive, negative, signed) x (small, big);
                                     chain form: up to twelve, each one 16-bit chain or a set of single columns (group_kind) */
#define CRB_CTRL_BYTES 512          /* head of the tiled kernel's shared memory: ring barriers and tile descriptors */
#define CRB_DIRECT_THREADS 256    /* block size (= frames per tile) of the direct kernel */
#ifndef CRB_RING_STAGES
#define CRB_RING_STAGES 2           /* measured: 4 CTAs x 2 stages beats 3 CTAs x 3 stages */
#endif

/* A run of columns of the per-phase table that share a weight sign and read consecutive input
   frames: columns [col, col+len) multiply frames [off, off+len) after the window start. */
typedef struct crb_run {
	int32_t col, len, off;
	int16_t negative;            /* 0: weights > 0 in every phase row, columns hold |k|;  1: weights < 0 in every row, columns hold |k| and
	                                the chain is subtracted at the end;  2: the weight changes sign between rows (a zero crossing of the
	                                stretched kernel moves across this tap), the column holds the signed weight */
	int16_t big;                 /* 1: columns hold |k| (up to 65536), multiplicand is sample << 16;
	                                0: columns hold |k| << 16 (|k| < 32768), multiplicand is the sample */
} crb_run;

/* Everything the kernels need to know about one (table, configuration, increment, channels).
   Passed by value as a __grid_constant__ kernel parameter. */
typedef struct crb_geometry {
	uint32_t channels;
	uint32_t increment;          /* 16.16, H:647 */
	uint32_t step;               /* kernel_step_size, H:637 */
	uint32_t delta;              /* stretched_kernel_radius_delta, H:636 */
	uint32_t radius_int;         /* integer_stretched_kernel_radius, H:635 */
	uint32_t radius_fx;          /* stretched_kernel_radius (fits: < 3 * 2^28) */
	uint32_t ks0;                /* table index of the first tap at e = 0 */
	uint32_t n_breaks;
	uint32_t breaks[CRB_MAX_BREAKS]; /* e thresholds where the tap count changes inside one ks value */
	uint32_t n_rows;
	uint32_t n_cols;             /* columns per row (taps, mixed-sign taps counted twice) */
	uint32_t row_words;          /* n_cols + 1 (reciprocal), padded to a multiple of 4 */
	uint32_t taps_max;           /* frames read after the window start */
	uint32_t n_runs;
	crb_run runs[CRB_MAX_RUNS];
	/* general kernel: columns are ordered by group = negative * 2 + big (negative as in crb_run: 0, 1, 2), every group
	   padded to an even number of columns (zero weight); groups[g] = {first column, columns}.  After the rows the table
	   holds one word per column: the byte offset of the input frame that column multiplies, relative to the window start. */
	uint32_t groups[CRB_GROUPS][2];
	uint32_t colinfo_words;      /* 0 for the packed unstretched table */
	uint32_t tile_out;           /* output frames per tile */
	uint32_t tile_in_frames;     /* frames of shared memory per stage */
	uint32_t stage_bytes;        /* bytes per stage, multiple of 16 */
	uint32_t unstretched5;       /* 1: step 1024, delta 0, five columns with signs + - + + - */
	uint32_t lane_stride;        /* odd s: consumer thread t takes frame (t * s) mod NT of every NT-frame block (NT = CRB_NT), chosen per plan so
	                                that the lanes of one shared-memory load hit different banks (1 = consecutive frames) */
	/* column rotation (general kernel): lane l of a warp starts every rotating group at pair ((rot * l) >> rot_shift) & rot_mask, so
	   that lanes whose frames are a multiple of 32 banks apart (integer down-sampling ratios) read different
	   columns -- different banks -- in the same instruction.  A rotating group is followed by a copy of its first
	   2 * rot_mask columns (no wrap-around test in the loop).  rot == 0: off. */
	uint32_t rot, rot_shift, rot_mask;
	uint32_t group_rot[CRB_GROUPS];       /* 0xFFFFFFFF when the group rotates (more than rot_mask pairs), else 0 */
	/* general kernel without rotation: the columns' frame offsets (bytes) also travel here, in the kernel parameters, so
	   that the tap loop reads them through the constant cache into uniform registers (no shared-memory load, no
	   address add per tap); const_offsets == 0: too many columns, or a rotating plan (per-lane columns) */
	uint32_t const_offsets;
	uint16_t col_off16[CRB_CONST_COLS];
	uint32_t small_taps;         /* 0, or 6 / 8 / 10 / 12: slightly stretched kernel (down-sampling by less than about 2, up to eight
	                                channels): rows hold that many SIGNED weights in tap order, then the reciprocal word, padded to a
	                                multiple of four words; the kernel is unrolled over the taps (crb_device.cu frame_sk) */
	uint32_t norm_mode;          /* last row word: 3, 2 = (recip - 32768) << 17, 1 = (recip - 32768) << 16, 0 = recip (see normalise()) */
	uint32_t n_stages;           /* depth of the input-window ring in shared memory */
	/* general kernel, chain form (chain_mode == 1; crb_kernels.cuh frame_chains): every column holds the plain weight (|k| for the
	   positive and negative classes, signed k for the signed class, |k| <= 65536) and the multiply-accumulate runs in full-rate
	   instructions.  group_kind[g] = class * 2 + single: class 0 / 1 / 2 = positive / negative / signed; single == 0: the group is
	   ONE chain whose |k| sum to at most 65535 in every phase row (its running sum lives in the upper 16 bits of a register and is
	   folded into the 32-bit accumulator once, at the end of the group); single == 1: every column is folded by itself. */
	uint32_t chain_mode, n_groups;
	uint8_t group_kind[CRB_GROUPS];
	uint32_t lock_slot_bytes[3]; /* unstretched kernel: bytes of a stage each stream of a lockstep job of 1, 2, 4 streams gets (index log2);
	                                0 = that many lockstep streams do not fit */
} crb_geometry;

/* A unit of work as the device sees it.  q0 = (position << 16 | fraction) + delta, i.e. the
   16.16 position of output frame 0 shifted by the radius delta, so that the first frame a
   window reads is ceil(q / 65536) (H:993, H:995). */
#define CRB_MAX_LOCKSTEP 4
typedef struct crb_device_job {
	const int16_t *in;           /* padded input, device */
	void *out;                   /* device */
	uint64_t q0;
	uint64_t first_out;          /* index of the first output frame of this job */
	uint64_t n_out;
	uint64_t in_frames;          /* padded frames available at `in` (total + 2R) */
	uint64_t tile_base;          /* exclusive prefix sum of tiles over jobs */
	uint64_t increment;          /* 16.16 step of this job; 0 = the plan's.  The phase table depends on the kernel
	                                geometry only, so jobs with different increments (pitch-bent voices) share a
	                                launch as long as none exceeds the plan's increment (tile sizing) */
	/* Lockstep streams: n_more (0, 1 or 3) further streams that share q0's fraction, first_out, n_out and increment, i.e.
	   the same phase for every output frame.  A thread then fetches a frame's phase row once and computes that
	   frame of all 1 + n_more streams with it (unstretched kernel only).  Such a job's tiles hold
	   tile_out / (1 + n_more) frames of each stream. */
	uint32_t n_more;
	uint32_t reserved;
	const int16_t *in_more[CRB_MAX_LOCKSTEP - 1];
	void *out_more[CRB_MAX_LOCKSTEP - 1];
	uint64_t in_frames_more[CRB_MAX_LOCKSTEP - 1];
} crb_device_job;

struct ClownResamplerB200_Plan {
	crb_geometry geo;
	uint64_t table_hash;
	unsigned table_id;           /* cached plans: which registered table (crb_api.c table_id) the plan was built from */
	uint32_t cfg_radius_fx, cfg_radius_int, cfg_delta, cfg_step;
	int32_t *host_rows;          /* n_rows * row_words */
	int32_t *host_table;         /* CRB_TABLE_SIZE, int32 copy of the caller's table */
	void *dev_rows;
	void *dev_table;
	int kernel_kind;             /* 0 tiled, 1 direct */
	uint32_t smem_bytes;
	double mean_taps;
	int device;
	int refcount;
	const void *launch_fn[3];    /* per output format: kernel instantiation the cached launch configuration belongs to */
	int blocks_per_sm[3];
};

/* ---- crb_plan.c ---- */
uint64_t crb_hash_table(const long *table);
int crb_plan_build_host(struct ClownResamplerB200_Plan *plan, const long *table,
	uint64_t radius_fx, uint64_t radius_int, uint64_t delta, uint64_t step,
	uint64_t increment, unsigned channels, uint32_t smem_budget_bytes);
void crb_set_error(const char *fmt, ...);

/* ---- crb_device.cu ---- */
int crb_dev_init(int device, int make_default);   /* device < 0: the default device; returns the device index or < 0 */
int crb_dev_default(void);
int crb_dev_count(void);
uint32_t crb_dev_smem_optin(int device);
int crb_dev_sm_count(int device);
int crb_dev_push(int device);                      /* make `device` current in this thread; returns the previous one for crb_dev_pop */
void crb_dev_pop(int previous);
void *crb_dev_alloc(size_t bytes);
void crb_dev_free(void *p);
void *crb_dev_pinned_alloc(size_t bytes);
void crb_dev_pinned_free(void *p);
int crb_dev_h2d(void *dst, const void *src, size_t bytes, void *stream);
int crb_dev_d2h(void *dst, const void *src, size_t bytes, void *stream);
int crb_dev_sync(void *stream);
int crb_dev_is_pinned(const void *host_pointer, size_t bytes);
int crb_dev_of_pointer(const void *device_pointer);
void *crb_dev_stream_create(void);
void crb_dev_stream_destroy(void *stream);
void *crb_dev_event_create(void);
void crb_dev_event_destroy(void *event);
int crb_dev_event_record(void *event, void *stream);
int crb_dev_event_sync(void *event);
int crb_dev_stream_wait_event(void *stream, void *event);
int crb_dev_plan_upload(struct ClownResamplerB200_Plan *plan);
void crb_dev_plan_release(struct ClownResamplerB200_Plan *plan);
/* jobs: host array; copied into kernel parameters (<= CRB_INLINE_JOBS) or a device array. */
int crb_dev_launch(struct ClownResamplerB200_Plan *plan, const crb_device_job *jobs, size_t n_jobs,
	uint64_t total_tiles, int out_format, void *stream);
/* same, with the job table already in device memory (uploaded by the caller, e.g. together with the input) */
int crb_dev_launch_resident(struct ClownResamplerB200_Plan *plan, const crb_device_job *device_jobs, size_t n_jobs,
	uint64_t total_tiles, int out_format, void *stream);
int crb_dev_fill_noise(int16_t *dst, uint32_t seed, uint32_t stream_id, uint64_t first_frame,
	uint64_t n_frames, uint32_t channels, void *stream);
int crb_dev_interleave(void *const *planes, void *interleaved, uint64_t frames, uint32_t channels, int word_bytes, int to_planes, void *stream);
int crb_dev_checksum(const void *src, uint64_t words, int word_bytes, unsigned long long *result, void *stream);

#ifdef __cplusplus
}
#endif
#endif
